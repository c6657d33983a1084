"""CPU tests of the host logic and of the C-ABI surface (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from tests.parity_util import ROOT, SHIPPED_KW, have_checkpoint, real_state_dict, random_state_dict, load_pairs
from diffphore_b200 import irreps as ir
from diffphore_b200.graph import collate, uncollate, DataLoader, graph_to_arrays, graph_from_arrays
from diffphore_b200.synthetic import make_pair, make_pairs


def test_library_loads_and_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, 'include', 'diffphore_b200.h')).read()
    declared = set(re.findall(r'\b(dp_[a-z0-9_]+)\s*\(', header))
    assert declared == set(built_lib.EXPORTS), declared ^ set(built_lib.EXPORTS)
    handle = ctypes.CDLL(built_lib.LIB_PATH)
    for name in declared:
        assert getattr(handle, name) is not None
    lib = built_lib.load()
    assert lib.dp_version() == 100


def test_no_cpu_fallback_when_library_missing(monkeypatch, built_lib):
    monkeypatch.setattr(built_lib, '_lib', None)
    monkeypatch.setattr(built_lib, 'LIB_PATH', '/nonexistent/libdiffphore_sm100.so')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        built_lib.load()


def test_product_does_not_import_oracle():
    for d in ('diffphore_b200', 'src'):
        for root, _, files in os.walk(os.path.join(ROOT, d)):
            for f in files:
                if f.endswith('.py'):
                    txt = open(os.path.join(root, f)).read()
                    assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, re.M), os.path.join(root, f)


def test_instruction_tables_match_generated_header():
    hdr = open(os.path.join(ROOT, 'diffphore_b200', 'csrc', 'tp_tables.cuh')).read()
    seq = [ir.parse_irreps(s) for s in ir.IRREP_SEQ(20, 10)]
    sh = ir.sh_irreps(2)
    for name, a, b, numel in [('TpL0', seq[0], seq[1], 600), ('TpL1', seq[1], seq[2], 1100), ('TpL2', seq[2], seq[3], 1600),
                              ('TpL3', seq[3], seq[3], 2200)]:
        assert ir.fctp_instructions(a, sh, b)[1] == numel
        assert re.search(rf'struct {name} \{{\s+static constexpr int D_IN = {ir.irreps_dim(a)}, D_OUT = {ir.irreps_dim(b)}, W = {numel},', hdr)


def test_reference_facing_model_strict_load_and_flag_guard():
    from models.score_model_phore import TensorProductScoreModel
    m = TensorProductScoreModel(None, torch.device('cpu'), None, **SHIPPED_KW)
    assert len(m.state_dict()) == 385
    sd = real_state_dict() if have_checkpoint() else random_state_dict(0)
    assert len(sd) == 385
    m.load_state_dict(sd, strict=True)
    with pytest.raises(NotImplementedError):
        TensorProductScoreModel(None, torch.device('cpu'), None, **dict(SHIPPED_KW, use_att=True))
    with pytest.raises(NotImplementedError):
        TensorProductScoreModel(None, torch.device('cpu'), None, **dict(SHIPPED_KW, atom_weight='softmax'))


def test_collate_uncollate_roundtrip_and_loader():
    gs = make_pairs(3, 10, 5)
    b = collate([g.clone() for g in gs])
    assert b.num_graphs == 3 and b['ligand'].pos.shape[0] == 30 and b['ligand'].batch.tolist() == [0] * 10 + [1] * 10 + [2] * 10
    assert int(b['ligand', 'ligand'].edge_index[:, 18:36].min()) >= 10
    back = uncollate(b)
    for g, h in zip(gs, back):
        assert torch.equal(g['ligand'].pos, h['ligand'].pos) and torch.equal(g['ligand', 'ligand'].edge_index, h['ligand', 'ligand'].edge_index)
        assert torch.equal(g['phore', 'phore'].edge_index, h['phore', 'phore'].edge_index)
        assert np.array_equal(g['ligand'].mask_rotate, h['ligand'].mask_rotate)
    assert [x.num_graphs for x in DataLoader(gs, batch_size=2)] == [2, 1]
    a = graph_to_arrays(gs[0], 'p_')
    g2 = graph_from_arrays(a, 'p_')
    assert torch.equal(g2['ligand'].x, gs[0]['ligand'].x) and torch.equal(g2['phore'].phoretype, gs[0]['phore'].phoretype)


def test_synthetic_generator_schema():
    g = make_pair(5, 32, 8)
    assert torch.equal(make_pair(5, 32, 8)['ligand'].pos, g['ligand'].pos)          # seeded
    lig, ph = g['ligand'], g['phore']
    ei = g['ligand', 'ligand'].edge_index
    assert lig.x.shape == (32, 16) and lig.norm.shape == (32, 33) and lig.phorefp.shape == (32, 11)
    assert ei.shape == (2, 62) and torch.equal(ei[:, 0::2], ei[:, 1::2].flip(0))
    d = (lig.pos[ei[0]] - lig.pos[ei[1]]).norm(dim=1)
    assert torch.allclose(d, torch.full_like(d, 1.5), atol=1e-4)
    rot = ei[:, lig.edge_mask]
    assert lig.mask_rotate.shape == (rot.shape[1], 32)
    for k in range(rot.shape[1]):
        assert not lig.mask_rotate[k, rot[0, k]] and lig.mask_rotate[k, rot[1, k]]      # torsion.py:89-90
    assert ph.x.shape == (8, 5) and ph.phoretype.shape == (8, 11) and float(ph.phoretype[:, 10].sum()) >= 1
    assert torch.allclose(ph.pos.mean(0), torch.zeros(3), atol=1e-5)
    dd = torch.cdist(lig.pos, lig.pos) + torch.eye(32) * 10
    assert float(dd.min()) > 1.49


def test_packing_layout():
    from diffphore_b200.engine import ModelWeights, PackedBatch
    w = ModelWeights(random_state_dict(0), 'cpu')
    gs = [make_pair(0, 10, 5), make_pair(1, 14, 6)]
    b = PackedBatch(gs, 3, w, torch.device('cpu'))
    assert b.B == 6 and b.n_lig == 3 * 10 + 3 * 14 and b.n_ph == 3 * 5 + 3 * 6
    assert b.lig_ptr.tolist() == [0, 10, 20, 30, 44, 58, 72]
    assert b.n_cross == 3 * 50 + 3 * 84 and b.cross_seg_lig[-1] == b.n_cross and b.cross_seg_ph[-1] == b.n_cross
    # canonical cross order is (lig, phore); the transposed list enumerates the same edges phore-major
    cl, cp = b.cross_lig.long(), b.cross_ph.long()
    assert torch.equal(cl[b.cross_perm_t.long()], b.cross_lig_t.long()) and torch.equal(cp[b.cross_perm_t.long()], b.cross_ph_t.long())
    assert sorted(b.cross_perm_t.tolist()) == list(range(b.n_cross))
    seg = b.cross_seg_ph.long()
    for q in (0, 7, b.n_ph - 1):
        assert set(b.cross_ph_t[seg[q]:seg[q + 1]].tolist()) == {q}
    # bond CSR
    bp = b.bond_ptr.long()
    ei = gs[1]['ligand', 'ligand'].edge_index
    a0 = 30
    for a in range(14):
        got = sorted((b.bond_dst[bp[a0 + a]:bp[a0 + a + 1]] - a0).tolist())
        assert got == sorted(ei[1][ei[0] == a].tolist())
    assert b.mask.numel() == 3 * gs[0]['ligand'].mask_rotate.size + 3 * gs[1]['ligand'].mask_rotate.size
    assert b.ll_cap >= b.n_bond and b.lig_static.shape == (b.n_lig, 20)


def test_step_constants_fold_the_sigma_embedding():
    from diffphore_b200.engine import ModelWeights, sinusoidal_embedding
    from diffphore_b200.tables import So3ScoreNorm, TorusScoreNorm
    sd = random_state_dict(0)
    w = ModelWeights(sd, 'cpu')
    sc = w.step_consts(0.35, So3ScoreNorm(), TorusScoreNorm(), dt=0.05)
    semb = sinusoidal_embedding(0.35)
    ref = sd['encoder.lig_edge_embedding.0.weight'][:, 4:24] @ semb + sd['encoder.lig_edge_embedding.0.bias']
    assert torch.allclose(sc[60:80], ref, atol=1e-6)
    tr_s = 0.1 ** 0.65 * 5.0 ** 0.35
    assert abs(float(sc[180]) * tr_s - 1) < 1e-5
    g = tr_s * np.sqrt(2 * np.log(50.0))
    assert abs(float(sc[183]) - g * g * 0.05) < 1e-6 and abs(float(sc[184]) - g * np.sqrt(0.05)) < 1e-6
    from utils.diffusion_utils import get_t_schedule, sinusoidal_embedding as se2
    assert torch.allclose(se2(torch.tensor([10000 * 0.35]), 20)[0], semb)
    assert np.allclose(get_t_schedule(20), np.linspace(1, 0, 21)[:-1])


# ---------------------------------------------------------------------------------------------------------------
# host side of the fused convolution (dp_conv_fused): tiles, operand images, device-side sample expansion
# ---------------------------------------------------------------------------------------------------------------
def test_greedy_tiles_rule():
    from diffphore_b200.engine import greedy_tiles
    assert greedy_tiles([8] * 64) == [0, 32]                                   # 32 nodes x 8 edges fill a 256-edge pair tile
    assert greedy_tiles([32] * 16) == [0, 8]
    assert greedy_tiles([200, 0, 0, 56, 1]) == [0, 4]                          # zero-degree nodes ride along, 257th edge opens a tile
    assert greedy_tiles([256, 256]) == [0, 1]
    assert greedy_tiles([79] * 7) == [0, 3, 6]                                 # 3 x 79 = 237 of 256 rows (79-point pharmacophore)
    assert greedy_tiles([0, 0, 257]) is None                                   # a node with more than 256 edges: unfused kernels
    assert greedy_tiles([0] * 600) == [0, 256, 512]                            # node cap: node_seg[] of the kernel holds 257 entries
    assert greedy_tiles([]) == []
    rng = np.random.default_rng(0)
    deg = rng.integers(0, 60, 500)
    t = greedy_tiles(deg) + [len(deg)]
    seg = np.concatenate([[0], np.cumsum(deg)])
    fill = [seg[b] - seg[a] for a, b in zip(t[:-1], t[1:])]
    assert max(fill) <= 256 and all(f + deg[b] > 256 for f, b in zip(fill[:-1], t[1:-1]))   # greedy: the next node would not fit


def test_grouped_tiles_equal_the_greedy_rule_restarted_per_group():
    """engine.grouped_tiles (static edge sets; vectorised across groups) = greedy_tiles applied to every group of consecutive
    graphs, the rule dp_build_tiles applies on the device to the dynamic edge sets."""
    from diffphore_b200.engine import grouped_tiles, greedy_tiles
    rng = np.random.default_rng(0)
    for _ in range(40):
        B = int(rng.integers(1, 30))
        npg = rng.integers(0, 12, size=B)
        deg = rng.integers(0, 60, size=int(npg.sum()))
        G = int(rng.integers(1, 9))
        off = np.concatenate([[0], np.cumsum(npg)])
        exp = [off[g0] + t for g0 in range(0, B, G) for t in greedy_tiles(deg[off[g0]:off[min(g0 + G, B)]])]
        assert list(grouped_tiles(deg, npg, group=G)) == exp
    assert grouped_tiles([3, 257], [2]) is None
    assert list(grouped_tiles([0] * 600, [600])) == [0, 256, 512]
    # 8-point pharmacophores with 24 edges: one tile per graph when restarted per graph, one per group of 8 (192 edges) now
    assert len(grouped_tiles([3] * 8 * 16, [8] * 16, group=8)) == 2


def test_fused_operand_images_reconstruct_the_weights():
    from diffphore_b200.engine import _make_w1img, _make_w2img112
    g = torch.Generator().manual_seed(0)
    w3, b3 = torch.randn(1100, 60, generator=g) * 3, torch.randn(1100, generator=g)
    img, inv = _make_w2img112(w3, b3)
    h = img.view(torch.float16).reshape(11, 2, 8, 14, 8, 8)                     # [chunk][hi|lo][k/8][n/8][n%8][k%8]
    x = (h[:, 0].double() + h[:, 1].double()).permute(0, 2, 3, 1, 4).reshape(11, 112, 64) * inv
    assert float((x[:, :100, :60].reshape(1100, 60) - w3).abs().max()) < 2e-6 * float(w3.abs().max())
    assert float((x[:, :100, 60].reshape(1100) - b3).abs().max()) < 2e-6 * float(w3.abs().max())
    assert float(x[:, 100:].abs().max()) == 0 and float(x[:, :, 61:].abs().max()) == 0
    w1, b1 = torch.randn(60, 60, generator=g), torch.randn(60, generator=g)
    img, inv = _make_w1img(w1, b1)
    h = img.view(torch.float16).reshape(2, 8, 8, 8, 8)                          # [hi|lo][k/8][n/8][n%8][k%8]
    x = (h[0].double() + h[1].double()).permute(1, 2, 0, 3).reshape(64, 64) * inv
    assert float((x[:60, :60] - w1).abs().max()) < 2e-6 * float(w1.abs().max())
    assert float((x[:60, 60] - b1).abs().max()) < 2e-6 * float(w1.abs().max())
    assert abs(float(x[60, 60]) - 1.0) < 1e-7 and float(x[61:].abs().max()) == 0   # constant-1 column passes through layer 1
    assert np.log2(inv) == round(np.log2(inv))                                  # exact power-of-two scale


def test_flat_weight_image_reconstructs_the_weights():
    """engine._make_w2imgflat (experimental layout of dp_conv_fused_flat): chunk c, column n of the image = weight column 112 c + n,
    zero beyond W; same power-of-two scale as the path-aligned image."""
    from diffphore_b200.engine import _make_w2imgflat, _make_w2img112
    g = torch.Generator().manual_seed(0)
    for W in (600, 1100, 1600, 2200):
        w3, b3 = torch.randn(W, 60, generator=g) / 8, torch.randn(W, generator=g)
        img, inv = _make_w2imgflat(w3, b3)
        nch = (W + 111) // 112
        h = img.view(torch.float16).reshape(nch, 2, 8, 14, 8, 8)                 # [c][hi|lo][k/8][n/8][n%8][k%8]
        x = (h[:, 0].double() + h[:, 1].double()).permute(0, 2, 3, 1, 4).reshape(nch * 112, 64) * inv
        assert float((x[:W, :60] - w3).abs().max()) < 2e-6 * float(w3.abs().max())
        assert float((x[:W, 60] - b3).abs().max()) < 2e-6 * float(b3.abs().max())
        assert float(x[W:].abs().max()) == 0 and float(x[:, 61:].abs().max()) == 0
        assert inv == _make_w2img112(w3, b3)[1] and img.numel() == nch * 28672


def test_packed_batch_expands_samples_like_a_naive_replication():
    """PackedBatch uploads per-pair arrays and expands the samples on the device; compare with collating deep copies."""
    from diffphore_b200.engine import ModelWeights, PackedBatch
    graphs = load_pairs('synthetic', 3, 12, 5) + load_pairs('synthetic', 1, 4, 4) + load_pairs('synthetic', 1, 3, 4)
    S = 3
    w = ModelWeights(random_state_dict(0), 'cpu')
    a = PackedBatch(graphs, S, w, 'cpu')
    b = PackedBatch([g for g in graphs for _ in range(S)], 1, w, 'cpu')          # every (pair, sample) as its own "pair"
    assert a.h2d_bytes < b.h2d_bytes
    for k, v in vars(b).items():
        va = getattr(a, k)
        if torch.is_tensor(v):
            assert va.dtype == v.dtype and torch.equal(va, v), k
        elif isinstance(v, tuple) and len(v) == 3 and torch.is_tensor(v[0]):
            assert torch.equal(va[0], v[0]) and va[2] == v[2], k
        elif isinstance(v, np.ndarray):
            assert np.array_equal(va, v), k
        elif k not in ('h2d_bytes', 'S', 'device'):
            assert va == v, k
    assert a.tiles_cross_lig[2] > 0 and a.tiles_cross_lig[0].shape[0] == a.tiles_cross_lig[2] + 1


def test_host_mask_expansion_equals_the_generic_level_expansion():
    """Host-packed batches build the expanded mask_rotate rows with S copies per pair (engine.expand_mask_rows_host); the result
    must equal the generic index expansion (pairs with and without rotatable bonds, several samples)."""
    from diffphore_b200.engine import ModelWeights, PackedBatch, expand_mask_rows_host, _pair_arrays
    graphs = load_pairs('synthetic', 3, 12, 5) + load_pairs('synthetic', 1, 3, 4) + load_pairs('synthetic', 2, 20, 6)
    w = ModelWeights(random_state_dict(0), 'cpu')
    for S in (1, 4):
        a = PackedBatch(graphs, S, w, 'cpu')
        mask, off = expand_mask_rows_host([_pair_arrays(g) for g in graphs], S)
        assert mask.dtype == a.mask.dtype and torch.equal(mask, a.mask)
        assert off.dtype == a.mask_off.dtype and torch.equal(off, a.mask_off)
    g0 = load_pairs('synthetic', 1, 3, 4)                                         # no rotatable bond anywhere
    a = PackedBatch(g0, 2, w, 'cpu')
    mask, off = expand_mask_rows_host([_pair_arrays(g) for g in g0], 2)
    assert mask.numel() == a.mask.numel() == 0 and torch.equal(off, a.mask_off)


@pytest.mark.parametrize('layout', ['paths', 'flat', 'flat_trim', 'gen2'])
def test_engine_conv_dispatch_follows_the_weight_layout(monkeypatch, layout):
    """Host glue of Engine._conv (no GPU: the library object is replaced by a recorder): DIFFPHORE_W2=paths calls dp_conv_fused with
    the path-aligned image; flat / flat_trim (the default) call dp_conv_fused_flat with the flat image and mode bit 4 for the
    trimmed last chunk; the second-generation kernel (default for layer 0 and the torsion convolution, DIFFPHORE_CONV_GEN=2 for
    all layers) gets the 96-column flat image."""
    from types import SimpleNamespace
    from diffphore_b200.engine import ModelWeights, Engine
    monkeypatch.setenv('DIFFPHORE_W2', 'flat_trim' if layout == 'gen2' else layout)
    monkeypatch.setenv('DIFFPHORE_CONV_GEN', '2' if layout == 'gen2' else 'auto')
    w = ModelWeights(random_state_dict(0), 'cpu')
    assert w.convs[('lig', 0)].gen == 2 and w.convs['tor'].gen == 2 and w.convs[('lig', 1)].gen == (2 if layout == 'gen2' else 1)
    eng = Engine(w)
    calls = []

    def rec(name):
        def f(*a):
            calls.append((name, a))
            return 0
        return f
    eng.lib = SimpleNamespace(dp_conv_fused=rec('dp_conv_fused'), dp_conv_fused_flat=rec('dp_conv_fused_flat'),
                              dp_conv_fused2=rec('dp_conv_fused2'))
    cw = w.convs[('lig', 3)]
    assert (getattr(cw, 'w2imgflat', None) is None) == (layout in ('paths', 'gen2'))
    z = lambda *s: torch.zeros(*s)
    zi = lambda *s: torch.zeros(*s, dtype=torch.int32)
    ws = SimpleNamespace(n_launches=0)
    tiles = (zi(2), None, 1)
    eng._conv(cw, ws, z(4, 20), None, z(3, 100), zi(4), z(3, 100), zi(4), None, None, 4, z(3, 100), zi(4), z(4, 9), 9, zi(4),
              z(3, 100), z(3, 100), 100, 1, 3, 0, 'lig3', tiles)
    (name, a), = calls
    assert ws.n_launches == 1
    assert name == {'paths': 'dp_conv_fused', 'gen2': 'dp_conv_fused2'}.get(layout, 'dp_conv_fused_flat')
    img = {'paths': cw.w2img112, 'gen2': getattr(cw, 'w2img96', None)}.get(layout, getattr(cw, 'w2imgflat', None))
    as_int = lambda v: v if isinstance(v, int) or v is None else v.value
    assert as_int(a[12]) == img.data_ptr() and a[0] == cw.layer_id
    assert a[27] == (1 | 16 if layout == 'flat_trim' else 1)                      # mode (+ trim bit)
    if layout == 'gen2':
        assert img.numel() == 23 * 24576
        return
    assert img.numel() == (22 if layout == 'paths' else 20) * 28672
