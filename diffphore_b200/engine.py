"""Host side of the B200 denoising path: weight folding, batch packing, buffer ownership and kernel sequencing.

PyTorch is used only for device memory, streams and one-off setup gathers; every per-step computation is a
hand-written sm_100a kernel reached through the C ABI (include/diffphore_b200.h).  The sequence mirrors
    TensorProductScoreModel.forward         /root/reference/src/models/score_model_phore.py:294-378
    LigPhoreEncoder.forward                 /root/reference/src/models/score_model_phore.py:644-712
    sampling_phore (one step)               /root/reference/src/utils/sampling.py:204-255
"""
import math
import os
from types import SimpleNamespace

import numpy as np
import torch

from . import lib as L
from .irreps import IRREP_SEQ, parse_irreps, irreps_dim, sh_irreps, fctp_instructions, full_tp_irreps_out

DEFAULT_CONFIG = dict(ns=20, nv=10, num_conv_layers=4, sigma_embed_dim=20, distance_embed_dim=20,
                      cross_distance_embed_dim=20, lig_max_radius=5.0, cross_max_distance=25.0,
                      center_max_distance=30.0, embedding_scale=10000, scaler=100.0,
                      clash_cutoff=[1.0, 2.0, 3.0, 4.0, 5.0], max_neighbors=32,
                      tr_sigma_min=0.1, tr_sigma_max=5.0, rot_sigma_min=0.1, rot_sigma_max=1.5,
                      tor_sigma_min=0.0314, tor_sigma_max=3.14, no_clamp=False)

LAYER_DIMS = [20, 50, 80, 100, 100]


def _f32(t):
    return t.detach().to(torch.float32).contiguous()


def sinusoidal_embedding(t, dim=20, scale=10000.0, max_positions=10000):
    """get_timestep_embedding('sinusoidal') of diffusion_utils.py:82-132 for one scalar t (fp32 like the reference)."""
    half = dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(max_positions) / (half - 1)))
    x = (torch.tensor([scale], dtype=torch.float32) * torch.tensor([t], dtype=torch.float32)).float()[:, None] * freq[None, :]
    return torch.cat([torch.sin(x), torch.cos(x)], dim=1)[0]


class ConvWeights:
    """One TensorProductConvLayer (smp:76-149) prepared for dp_edge_mlp / dp_tp_scatter."""

    @staticmethod
    def make_w2img(w3, b3):
        return _make_w2img(w3, b3)

    def __init__(self, sd, prefix, layer_id, in_irreps, sh_ir, out_irreps, device):
        instrs, numel = fctp_instructions(in_irreps, sh_ir, out_irreps)
        w3 = sd[prefix + '.fc.3.weight']
        if w3.shape[0] != numel:
            raise ValueError(f'{prefix}: fc.3 has {w3.shape[0]} rows, instruction table needs {numel}')
        self.layer_id, self.W = layer_id, numel
        self.in_dim, self.hid = sd[prefix + '.fc.0.weight'].shape[1], sd[prefix + '.fc.0.weight'].shape[0]
        self.w1 = _f32(sd[prefix + '.fc.0.weight']).to(device)
        self.b1 = _f32(sd[prefix + '.fc.0.bias']).to(device)
        self.w2t = torch.cat([_f32(w3).T, _f32(sd[prefix + '.fc.3.bias'])[None, :]], 0).contiguous().to(device)
        self.w2img, self.inv_wscale, self.w2img112 = None, 1.0, None
        if self.hid == 60 and self.in_dim == 60:
            img, self.inv_wscale = self.make_w2img(_f32(w3).cpu(), _f32(sd[prefix + '.fc.3.bias']).cpu())
            self.w2img = img.to(device)
            if numel % 100 == 0:
                img, s112 = _make_w2img112(_f32(w3).cpu(), _f32(sd[prefix + '.fc.3.bias']).cpu())
                assert s112 == self.inv_wscale
                self.w2img112 = img.to(device)
                # weight-column layout of dp_conv_fused: 'flat_trim' (default: consecutive 112-column chunks, last MMA trimmed;
                # measured +3.3 % on the cfg2 step, profiles/ab_flat_r2.txt), 'flat', or 'paths' (100-column path-aligned chunks)
                layout = os.environ.get('DIFFPHORE_W2', 'flat_trim')
                # generation of the fused kernel: 'auto' (default) = dp_conv_fused2 (csrc/conv_fused2.cuh: next pair tile's operands
                # prepared by a dedicated warpgroup) where it measured faster on the B200 - the short weight streams of layer 0 and the
                # torsion convolution - and dp_conv_fused elsewhere; '1' / '2' force one generation (results are identical bit for bit)
                gen = os.environ.get('DIFFPHORE_CONV_GEN', 'auto')
                self.gen = (2 if layer_id in (L.TP_L0, L.TP_TOR) else 1) if gen == 'auto' else int(gen)
                if self.gen == 2:
                    img, s96 = _make_w2imgflat(_f32(w3).cpu(), _f32(sd[prefix + '.fc.3.bias']).cpu(), 96)
                    assert s96 == self.inv_wscale
                    self.w2img96 = img.to(device)
                elif layout in ('flat', 'flat_trim'):
                    img, sflat = _make_w2imgflat(_f32(w3).cpu(), _f32(sd[prefix + '.fc.3.bias']).cpu())
                    assert sflat == self.inv_wscale
                    self.w2imgflat = img.to(device)
                    self.flat_mode_bits = 16 if layout == 'flat_trim' else 0      # bit 4 of `mode`: last chunk's MMA trimmed
                img, self.inv_w1scale = _make_w1img(_f32(sd[prefix + '.fc.0.weight']).cpu(), _f32(sd[prefix + '.fc.0.bias']).cpu())
                self.w1img = img.to(device)
        elif self.hid == 40 and self.in_dim == 40 and layer_id == L.TP_FINAL and os.environ.get('DIFFPHORE_FINAL', 'fused') == 'fused':
            # final_conv (fc 40 -> 40 -> 200, attributes [centre-edge embedding | atom scalars]) on the fused kernel: both MLP
            # layers zero padded to the kernel's 60 / 60 (the padded hidden units are ReLU(0) = 0; the kernel's third attribute
            # block re-reads the atom scalars against zero weights, so the operand scale - the row maximum - is unchanged)
            w1p, b1p = torch.zeros(60, 60), torch.zeros(60)
            w1p[:40, :40] = _f32(sd[prefix + '.fc.0.weight']).cpu()
            b1p[:40] = _f32(sd[prefix + '.fc.0.bias']).cpu()
            w3p = torch.zeros(numel, 60)
            w3p[:, :40] = _f32(w3).cpu()
            img, self.inv_wscale = _make_w2imgflat(w3p, _f32(sd[prefix + '.fc.3.bias']).cpu())
            self.w2imgflat = self.w2img112 = img.to(device)                  # (w2img112: the fused path's "images present" marker)
            self.flat_mode_bits, self.gen = 16, 1
            img, self.inv_w1scale = _make_w1img(w1p, b1p)
            self.w1img = img.to(device)
        # eval BatchNorm (e3nn.nn.BatchNorm) and path weights folded into per-component scale/shift
        bw, bb = sd[prefix + '.batch_norm.weight'].double(), sd[prefix + '.batch_norm.bias'].double()
        rm, rv = sd[prefix + '.batch_norm.running_mean'].double(), sd[prefix + '.batch_norm.running_var'].double()
        fan = {}
        for ins in instrs:
            fan[ins.io] = fan.get(ins.io, 0) + in_irreps[ins.i1][0]
        scale, shift, iw, ib = [], [], 0, 0
        for io, (mul, l, p) in enumerate(out_irreps):
            pw = math.sqrt((2 * l + 1) / fan[io]) if io in fan else 0.0
            s = bw[iw:iw + mul] / torch.sqrt(rv[iw:iw + mul] + 1e-5)
            if l == 0 and p == 1:
                sh = bb[ib:ib + mul] - rm[ib:ib + mul] * s
                ib += mul
            else:
                sh = torch.zeros(mul, dtype=torch.float64)
            scale.append((s * pw).repeat_interleave(2 * l + 1))
            shift.append(sh.repeat_interleave(2 * l + 1))
            iw += mul
        self.oscale = torch.cat(scale).float().contiguous().to(device)
        self.oshift = torch.cat(shift).float().contiguous().to(device)
        self.d_in, self.d_out = irreps_dim(in_irreps), irreps_dim(out_irreps)
        # channel-mixing FLOPs per edge of the tensor product: sum over paths 2 * mul_in * mul_out * (2 l_out + 1)
        self.tp_flops = sum(2 * in_irreps[i.i1][0] * out_irreps[i.io][0] * (2 * out_irreps[i.io][1] + 1) for i in instrs)


def _make_w2img(w3, b3):
    """Shared-memory image of the second-layer weights for dp_edge_mlp_tc.  W2aug[n, 0:60] = fc.3.weight, [n, 60] = bias,
    zero padded to K = 64 and to a multiple of 64 columns, multiplied by a power of two 2^k that brings max|W2aug| into
    [2^12, 2^13), split into fp16 hi = fp16(x) and lo = fp16(x - hi), stored per 64-column half-chunk as
    [hi|lo][k/8][n/8][n%8][k%8] (canonical K-major no-swizzle core-matrix layout of the tcgen05 smem descriptors).
    Returns (uint8 image tensor, 2^-k)."""
    W = w3.shape[0]
    nch = (W + 63) // 64
    x = torch.zeros(nch * 64, 64, dtype=torch.float32)
    x[:W, :60] = w3
    x[:W, 60] = b3
    m = float(x.abs().max())
    k = 12 - math.floor(math.log2(m)) if m > 0 and math.isfinite(m) else 0
    xs = x * (2.0 ** k)
    hi = xs.half()
    lo = (xs - hi.float()).half()
    img = torch.stack([hi, lo], 0).reshape(2, nch, 8, 8, 8, 8).permute(1, 0, 4, 2, 3, 5)       # [c][h][kc][ng][r][j]
    return img.contiguous().reshape(-1).view(torch.uint8), 2.0 ** (-k)


def _make_w2img112(w3, b3):
    """Shared-memory image of the second-layer weights for dp_conv_fused: like _make_w2img, but per 100-column chunk of
    the e3nn weight layout, zero padded to the MMA N of 112: [chunk][hi|lo][k/8][n/8 (14)][n%8][k%8] fp16.
    Returns (uint8 image tensor, 2^-k)."""
    W = w3.shape[0]
    assert W % 100 == 0
    nch = W // 100
    x = torch.zeros(nch, 112, 64, dtype=torch.float32)
    x[:, :100, :60] = w3.reshape(nch, 100, 60)
    x[:, :100, 60] = b3.reshape(nch, 100)
    m = float(x.abs().max())
    k = 12 - math.floor(math.log2(m)) if m > 0 and math.isfinite(m) else 0
    xs = x * (2.0 ** k)
    hi = xs.half()
    lo = (xs - hi.float()).half()
    img = torch.stack([hi, lo], 1).reshape(nch, 2, 14, 8, 8, 8).permute(0, 1, 4, 2, 3, 5)     # [c][h][kc][ng][r][j]
    return img.contiguous().reshape(-1).view(torch.uint8), 2.0 ** (-k)


def _make_w2imgflat(w3, b3, n=112):
    """Flat layout for dp_conv_fused_flat (n = 112) and dp_conv_fused2 (n = 96): like _make_w2img112, but the W columns are cut into
    consecutive n-column chunks regardless of the path boundaries (zero padding in the last chunk only):
    [ceil(W/n)][hi|lo][k/8][n/8][8][8] fp16.  Same power-of-two scale as _make_w2img112 (the maximum is the same).
    Returns (uint8 image tensor, 2^-k)."""
    W = w3.shape[0]
    nch = (W + n - 1) // n
    x = torch.zeros(nch * n, 64, dtype=torch.float32)
    x[:W, :60] = w3
    x[:W, 60] = b3
    x = x.reshape(nch, n, 64)
    m = float(x.abs().max())
    k = 12 - math.floor(math.log2(m)) if m > 0 and math.isfinite(m) else 0
    xs = x * (2.0 ** k)
    hi = xs.half()
    lo = (xs - hi.float()).half()
    img = torch.stack([hi, lo], 1).reshape(nch, 2, n // 8, 8, 8, 8).permute(0, 1, 4, 2, 3, 5)     # [c][h][kc][ng][r][j]
    return img.contiguous().reshape(-1).view(torch.uint8), 2.0 ** (-k)


def _make_w1img(w1, b1):
    """Shared-memory image of the first-layer weights for dp_conv_fused: W1aug[n, 0:60] = fc.0.weight, [n, 60] = fc.0.bias,
    W1aug[60, 60] = 1 (passes the constant-1 attribute column through ReLU into the bias column of the second layer), zero
    padded to 64 x 64, scaled by 2^k into [2^12, 2^13), fp16 hi | lo, [hi|lo][k/8][n/8][n%8][k%8].  Returns (image, 2^-k)."""
    x = torch.zeros(64, 64, dtype=torch.float32)
    x[:60, :60] = w1
    x[:60, 60] = b1
    x[60, 60] = 1.0
    m = float(x.abs().max())
    k = 12 - math.floor(math.log2(m)) if m > 0 and math.isfinite(m) else 0
    xs = x * (2.0 ** k)
    hi = xs.half()
    lo = (xs - hi.float()).half()
    img = torch.stack([hi, lo], 0).reshape(2, 8, 8, 8, 8).permute(0, 3, 1, 2, 4)                 # [h][kc][ng][r][j]
    return img.contiguous().reshape(-1).view(torch.uint8), 2.0 ** (-k)


# Consecutive graphs whose nodes share one run of tiles (the greedy rule restarts at every group, the half-empty last tile of
# a group is the cost).  'auto': ~1024 ligand atoms per group (32 graphs at 32 atoms, 8 at 128) - the device-side builder works
# a group per CTA and falls back to a sequential walk beyond 4096 nodes (conv_fused.cuh, TILE_WALK_CAP).
_TILE_GROUP_ENV = os.environ.get('DIFFPHORE_TILE_GROUP', 'auto')
TILE_GROUP = 8 if _TILE_GROUP_ENV == 'auto' else int(_TILE_GROUP_ENV)


def tile_group_for(n_atoms, n_graphs):
    if _TILE_GROUP_ENV != 'auto':
        return TILE_GROUP
    return int(min(64, max(1, 1024 // max(1, -(-n_atoms // max(1, n_graphs))))))


TILE_EDGES = 256    # edges (and nodes) per pair tile of dp_conv_fused: two M = 128 MMA operands


HOST_PACK_MAX_GRAPHS = 64    # batches of up to this many (pair, sample) graphs are expanded on the host (PackedBatch)


def grouped_tiles(deg, nodes_per_graph, group=TILE_GROUP, cap=TILE_EDGES):
    """greedy_tiles restarted at every group of `group` consecutive graphs instead of at every graph, for ALL groups at once.
    Like tile_walk in conv_fused.cuh: the tile that starts at node i ends before nxt(i) = the first j > i with
    seg[j + 1] - seg[i] > cap or j - i == cap (one vectorised searchsorted over all nodes), then every group hops from tile
    start to tile start (numpy, vectorised across groups: as many steps as the longest group has tiles).
    deg: edge count of every node of the expanded batch, nodes_per_graph: [B].  Returns the first node of every tile
    (global node indices, int64) or None if a node exceeds the cap."""
    deg = np.asarray(deg, dtype=np.int64)
    if deg.size and int(deg.max()) > cap:
        return None
    n = int(deg.size)
    gstart = np.concatenate([[0], np.cumsum(nodes_per_graph)])[::group].astype(np.int64)
    gend = np.append(gstart[1:], int(np.sum(nodes_per_graph))).astype(np.int64)
    if gstart.size and gstart[-1] >= n and gend[-1] <= gstart[-1]:
        gstart, gend = gstart[:-1], gend[:-1]                               # (a trailing empty group)
    seg = np.concatenate([[0], np.cumsum(deg)])
    idx = np.arange(n, dtype=np.int64)
    gid = np.searchsorted(gstart, idx, side='right') - 1
    j = np.searchsorted(seg, seg[:-1] + cap, side='right') - 1             # largest j with seg[j] - seg[i] <= cap
    nxt = np.maximum(idx + 1, np.minimum(np.minimum(j, idx + cap), gend[gid] if n else idx))
    out, cur = [], gstart[gstart < gend]
    end = gend[gstart < gend]
    while cur.size:
        out.append(cur)
        step = nxt[cur]
        keep = step < end
        cur, end = step[keep], end[keep]
    return np.sort(np.concatenate(out)) if out else np.zeros(0, np.int64)


def greedy_tiles(deg, cap=TILE_EDGES):
    """Node-aligned pair tiles for dp_conv_fused: runs of whole output nodes with <= cap edges and <= cap nodes (same rule
    as tile_walk in conv_fused.cuh).  deg: per-node edge counts of ONE graph (or group).  Returns the first node of every
    tile, or None if a single node exceeds the cap."""
    deg = np.asarray(deg, dtype=np.int64)
    if deg.size and deg.min() == deg.max():                    # uniform degrees (complete bipartite cross edges): closed form
        d = int(deg[0])
        return None if d > cap else list(range(0, len(deg), cap if d == 0 else min(cap // d, cap)))
    first, fill, nodes = [], 0, 0
    for n, d in enumerate(deg):
        d = int(d)
        if d > cap:
            return None
        if not first or fill + d > cap or nodes == cap:
            first.append(n)
            fill = nodes = 0
        fill += d
        nodes += 1
    return first


class ModelWeights:
    """Everything the kernels need from a reference-format state_dict (checkpoint loads unchanged)."""

    def __init__(self, state_dict, device, config=None):
        self.cfg = dict(DEFAULT_CONFIG)
        if config:
            self.cfg.update(config)
        self.device = device = torch.device(device)
        sd = {k: v.detach().cpu() for k, v in state_dict.items()}       # weight folding is host work (a model moved to CUDA hands CUDA tensors)
        self.lib = L.load()
        ns, nv = self.cfg['ns'], self.cfg['nv']
        if (ns, nv, self.cfg['num_conv_layers']) != (20, 10, 4):
            raise NotImplementedError('kernels are specialised for ns=20, nv=10, num_conv_layers=4 (shipped config)')
        hard = dict(sigma_embed_dim=20, distance_embed_dim=20, cross_distance_embed_dim=20, max_neighbors=32)
        bad = {k: self.cfg[k] for k, v in hard.items() if self.cfg[k] != v}
        if bad or len(self.cfg['clash_cutoff']) != 5:
            raise NotImplementedError(f'kernels hard-code {hard} and five clash cut-offs (shipped config); got {bad or self.cfg["clash_cutoff"]}')
        seq = [parse_irreps(s) for s in IRREP_SEQ(ns, nv)]
        sh = sh_irreps(2)
        sh45, _ = full_tp_irreps_out(sh, [(1, 2, 1)])
        self.convs = {}
        for l in range(4):
            for fam in ('lig', 'phore', 'lig_to_phore', 'phore_to_lig', 'lig_to_phore_norm', 'phore_to_lig_norm'):
                if l == 3 and fam in ('phore', 'lig_to_phore', 'lig_to_phore_norm'):
                    continue                                            # never executed (smp:691)
                self.convs[(fam, l)] = ConvWeights(sd, f'encoder.{fam}_conv_layers.{l}', l, seq[min(l, 3)], sh,
                                                   seq[min(l + 1, 3)], device)
        self.convs['final'] = ConvWeights(sd, 'final_conv', L.TP_FINAL, seq[3], sh, parse_irreps('2x1o + 2x1e'), device)
        self.convs['tor'] = ConvWeights(sd, 'tor_bond_conv', L.TP_TOR, seq[3], sh45, parse_irreps(f'{ns}x0o + {ns}x0e'),
                                        device)
        # small weights
        self._keep = []

        def dev(t):
            t = _f32(t).to(device)
            self._keep.append(t)
            return t.data_ptr()

        def mlp(prefix, bias=True):
            return L.DpMlp(dev(sd[prefix + '.0.weight']), dev(sd[prefix + '.0.bias']) if bias else None,
                           dev(sd[prefix + '.3.weight']), dev(sd[prefix + '.3.bias']) if bias else None)

        sw = L.DpSmallWeights()
        sw.lig_edge, sw.pp_edge = mlp('encoder.lig_edge_embedding'), mlp('encoder.phore_edge_embedding')
        sw.cross_edge = mlp('encoder.cross_edge_embedding')
        sw.cdt, sw.pdt = mlp('encoder.cross_distance_transition'), mlp('encoder.phore_direction_transition')
        sw.pmt = mlp('encoder.phoretype_match_transition')
        sw.center_edge, sw.final_edge = mlp('center_edge_embedding'), mlp('final_edge_embedding')
        sw.tr_final, sw.rot_final = mlp('tr_final_layer'), mlp('rot_final_layer')
        sw.boarder_tables = dev(torch.stack([sd[f'encoder.boarder_embedding.atom_embedding_list.{i}.weight'] for i in range(5)]))
        sw.boarder_w = dev(sd['encoder.boarder_embedding.linear.weight'][:, 0])
        sw.boarder_b = dev(sd['encoder.boarder_embedding.linear.bias'])
        sw.tor_w0 = dev(sd['tor_final_layer.0.weight'])
        sw.tor_w3 = dev(sd['tor_final_layer.3.weight'][0])
        self.sw = sw
        # host copies for the per-step folds and the static embeddings
        self.h = {k: _f32(v).cpu() for k, v in sd.items() if v.dim() <= 2 and v.numel() < 10000}
        self.lig_tables = [_f32(sd[f'encoder.lig_node_embedding.atom_embedding_list.{i}.weight']).to(device) for i in range(16)]
        self.ph_tables = [_f32(sd[f'encoder.phore_node_embedding.atom_embedding_list.{i}.weight']).to(device) for i in range(3)]
        self.ph_lin_w = _f32(sd['encoder.phore_node_embedding.linear.weight']).to(device)
        # constants
        c = L.DpConstants()
        for i, key in enumerate(['encoder.lig_distance_expansion.offset', 'encoder.phore_distance_expansion.offset',
                                 'encoder.cross_distance_expansion.offset', 'center_distance_expansion.offset']):
            off = _f32(sd[key]).cpu()
            for k in range(20):
                c.rbf_mu[i][k] = float(off[k])
            c.rbf_coeff[i] = -0.5 / (off[1] - off[0]).item() ** 2           # smp:1010
        for i, v in enumerate(self.cfg['clash_cutoff']):
            c.clash_cutoff[i] = float(v)
        c.lig_radius, c.scaler = float(self.cfg['lig_max_radius']), float(self.cfg['scaler'])
        c.max_neighbors, c.no_clamp = int(self.cfg['max_neighbors']), int(bool(self.cfg['no_clamp']))
        self.consts = c
        if device.type == 'cuda':           # host-only construction (CPU tests of the packing logic) skips the upload
            L.check(self.lib.dp_set_constants(c), 'dp_set_constants')

    # ------------------------------------------------------------------ per-noise-level constant block
    def t_to_sigma(self, t):
        """diffusion_utils.py:16-20, evaluated in fp32 like the reference's tensor path (complex_t is fp32)."""
        c = self.cfg
        tt = torch.tensor(t, dtype=torch.float32)
        return tuple(float((c[f'{k}_sigma_min'] ** (1 - tt) * c[f'{k}_sigma_max'] ** tt).item()) for k in ('tr', 'rot', 'tor'))

    def step_consts(self, t, so3_norm, torus_norm, dt=None, ode=False):
        """256-float block for noise level t.  so3_norm / torus_norm: callables sigma(np.float32 array) -> array
        (utils/so3.py:92-96, utils/torus.py:82-86).  dt: Euler–Maruyama step (None -> score model only).
        ode: probability-flow update 0.5 g^2 dt score without noise (sampling.py:226-228,240-241)."""
        c, h = self.cfg, self.h
        semb = sinusoidal_embedding(float(t), c['sigma_embed_dim'], c['embedding_scale'])
        out = torch.zeros(L.SC['SIZE'], dtype=torch.float32)
        out[0:20] = semb

        def fold(wkey, bkey, lo, hi):
            return h[wkey][:, lo:hi] @ semb + h[bkey]

        out[20:40] = fold('encoder.lig_node_embedding.linear.weight', 'encoder.lig_node_embedding.linear.bias', 0, 20)
        out[40:60] = fold('encoder.phore_node_embedding.linear.weight', 'encoder.phore_node_embedding.linear.bias', 2, 22)
        out[60:80] = fold('encoder.lig_edge_embedding.0.weight', 'encoder.lig_edge_embedding.0.bias', 4, 24)
        out[80:100] = fold('encoder.phore_edge_embedding.0.weight', 'encoder.phore_edge_embedding.0.bias', 0, 20)
        out[100:120] = fold('encoder.cross_edge_embedding.0.weight', 'encoder.cross_edge_embedding.0.bias', 0, 20)
        out[120:140] = fold('center_edge_embedding.0.weight', 'center_edge_embedding.0.bias', 20, 40)
        out[140:160] = fold('tr_final_layer.0.weight', 'tr_final_layer.0.bias', 1, 21)
        out[160:180] = fold('rot_final_layer.0.weight', 'rot_final_layer.0.bias', 1, 21)
        tr_s, rot_s, tor_s = self.t_to_sigma(t)
        out[180] = float(np.float32(1.0) / np.float32(tr_s))
        out[181] = float(np.asarray(so3_norm(np.asarray([rot_s], dtype=np.float32)))[0])
        out[182] = float(np.sqrt(np.float32(np.asarray(torus_norm(np.asarray([tor_s], dtype=np.float32)))[0])))
        if dt is not None:
            # sampling.py:223-246 — sigma from numpy float64 t, g in float64, cast to fp32 at the multiply
            t64 = float(t)
            trs = c['tr_sigma_min'] ** (1 - t64) * c['tr_sigma_max'] ** t64
            rots = c['rot_sigma_min'] ** (1 - t64) * c['rot_sigma_max'] ** t64
            tors = c['tor_sigma_min'] ** (1 - t64) * c['tor_sigma_max'] ** t64
            tr_g = trs * math.sqrt(2 * math.log(c['tr_sigma_max'] / c['tr_sigma_min']))
            rot_g = 2 * rots * math.sqrt(math.log(c['rot_sigma_max'] / c['rot_sigma_min']))
            tor_g = tors * math.sqrt(2 * math.log(c['tor_sigma_max'] / c['tor_sigma_min']))
            out[183], out[184] = tr_g ** 2 * dt, tr_g * math.sqrt(dt)
            out[185], out[186] = rot_g ** 2 * dt, rot_g * math.sqrt(dt)
            out[187], out[188] = tor_g ** 2 * dt, tor_g * math.sqrt(dt)
            if ode:
                out[183], out[185], out[187] = 0.5 * tr_g ** 2 * dt, 0.5 * rot_g ** 2 * dt, 0.5 * tor_g ** 2 * dt
                out[184] = out[186] = out[188] = 0.0
        return out


# =====================================================================================================================
# batch packing
# =====================================================================================================================
def _np(t, dt):
    """Tensor / array -> numpy of dtype dt without a copy when it already has it (host packing is on the e2e critical path)."""
    a = t.numpy() if isinstance(t, torch.Tensor) and t.device.type == 'cpu' else (t.cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t))
    return a if a.dtype == dt else a.astype(dt)


def _pair_arrays(g):
    """numpy view of one HeteroGraph / PyG HeteroData pair in the kernels' layout (local indices)."""
    lig, ph = g['ligand'], g['phore']
    n = lig.pos.shape[0]
    ei = _np(g['ligand', 'ligand'].edge_index, np.int64)
    ea = _np(g['ligand', 'ligand'].edge_attr, np.float32)
    btype = ea.argmax(1).astype(np.int32) if ea.size else np.zeros(0, np.int32)
    mask = _np(lig.edge_mask, np.bool_)
    order = np.argsort(ei[0], kind='stable')
    bond_ptr = np.zeros(n + 1, np.int32)
    np.cumsum(np.bincount(ei[0], minlength=n), out=bond_ptr[1:])
    mr = lig.mask_rotate
    mr = _np(mr if isinstance(mr, (np.ndarray, torch.Tensor)) else mr[0], np.uint8).reshape(int(mask.sum()), n)
    pe = _np(g['phore', 'phore'].edge_index, np.int64)
    po = np.argsort(pe[0], kind='stable')
    P = ph.pos.shape[0]
    pp_ptr = np.zeros(P + 1, np.int32)
    np.cumsum(np.bincount(pe[0], minlength=P), out=pp_ptr[1:])
    f32 = np.float32
    return SimpleNamespace(
        n=n, P=P, x=_np(lig.x, np.int64), pos=_np(lig.pos, f32), norm=_np(lig.norm, f32).reshape(n, 33),
        phorefp=_np(lig.phorefp, f32), na1=_np(lig.norm_angle1, f32), na2=_np(lig.norm_angle2, f32),
        bond_ptr=bond_ptr, bond_dst=ei[1][order].astype(np.int32), bond_type=btype[order],
        rot_u=ei[0][mask].astype(np.int32), rot_v=ei[1][mask].astype(np.int32), mask=mr,
        px=_np(ph.x, f32), ppos=_np(ph.pos, f32), pnorm=_np(ph.norm, f32), ptype=_np(ph.phoretype, f32),
        pp_src=pe[0][po].astype(np.int32), pp_dst=pe[1][po].astype(np.int32), pp_ptr=pp_ptr)


def expand_mask_rows_host(pa, S):
    """mask_rotate rows of the expanded batch (pair-major: the S samples of a pair carry the pair's rows) and the byte offset of
    every graph's rows, built on the host with S copies per pair; equal to the generic Level expansion (tested)."""
    m_g = np.repeat(np.asarray([len(q.rot_u) * q.n for q in pa], dtype=np.int64), S)
    mask = np.concatenate([np.tile(np.asarray(q.mask).reshape(-1).astype(np.uint8), S) for q in pa])
    off = np.concatenate([np.zeros(1, np.int64), np.cumsum(m_g)[:-1]])
    return torch.from_numpy(mask), torch.from_numpy(off)


class PackedBatch:
    """Flattened device arrays for B graphs (PyG-Batch-like; graph g = pair g // samples, sample g % samples).

    Only per-PAIR arrays cross the PCIe bus; the `samples` copies of every pair (identical topology and features, the
    reference builds them with copy.deepcopy, inference.py:196) are expanded on the device with index arithmetic."""

    def __init__(self, graphs, samples_per_graph, weights, device):
        S = samples_per_graph
        pa = [_pair_arrays(g) for g in graphs]
        Np = len(pa)
        B = Np * S
        self.B, self.S, self.device = B, S, torch.device(device)
        self.h2d_bytes = 0
        # Small batches (the reference-facing forward(data) / 4-40 sample jobs) are expanded on the HOST and uploaded as finished
        # arrays: the index arithmetic below is ~270 tiny ATen kernels on the device, which cost more than the whole score model
        # there.  Large batches keep the device-side expansion (only per-pair arrays cross PCIe).  Same arithmetic either way.
        final_device = self.device
        device = torch.device('cpu') if (final_device.type == 'cuda' and B <= HOST_PACK_MAX_GRAPHS) else final_device
        on_host = device != final_device

        def up(a):
            t = torch.from_numpy(np.ascontiguousarray(a))
            if on_host:
                return t
            self.h2d_bytes += t.numel() * t.element_size()
            if device.type == 'cuda':
                return t.pin_memory().to(device, non_blocking=True)
            return t.to(device)

        cat = lambda key, dt: up(np.concatenate([np.asarray(getattr(q, key)).reshape(-1) if np.asarray(getattr(q, key)).ndim == 1
                                                 else np.asarray(getattr(q, key)) for q in pa]).astype(dt))
        cnt = lambda vals: up(np.asarray(vals, dtype=np.int64))
        # ---- per-pair element counts (host) and their device copies
        n_p, P_p = np.asarray([q.n for q in pa]), np.asarray([q.P for q in pa])
        nrot_p, nb_p, npp_p = (np.asarray([len(q.rot_u) for q in pa]), np.asarray([len(q.bond_dst) for q in pa]),
                               np.asarray([len(q.pp_src) for q in pa]))
        t_lig = [greedy_tiles(np.full(q.n, q.P)) for q in pa]
        t_ph = [greedy_tiles(np.full(q.P, q.n)) for q in pa]
        t_pp = [greedy_tiles(np.diff(q.pp_ptr)) for q in pa]
        self.n_per, self.P_per, self.nrot_per = np.repeat(n_p, S), np.repeat(P_p, S), np.repeat(nrot_p, S)
        self.n_lig, self.n_ph, self.n_rot = int(n_p.sum()) * S, int(P_p.sum()) * S, int(nrot_p.sum()) * S
        self.n_bond, self.n_pp, self.n_cross = int(nb_p.sum()) * S, int(npp_p.sum()) * S, int((n_p * P_p).sum()) * S
        self.max_atoms, self.max_rot = int(n_p.max()), int(nrot_p.max())
        k = weights.cfg['max_neighbors']
        self.ll_cap = int(self.n_bond + S * np.sum(n_p * np.minimum(n_p - 1, k + 1)))
        self.tor_cap = int(S * np.sum(nrot_p * np.minimum(n_p, k)))
        i64 = dict(dtype=torch.int64, device=device)
        pair_of_graph = torch.arange(Np, **i64).repeat_interleave(S)

        def excl(c):                                    # exclusive prefix sum with the total appended: [len + 1]
            return torch.cat([torch.zeros(1, **i64), torch.cumsum(c, 0)])

        class Level:
            """Expansion of one per-pair element kind (atoms, bonds, ...) over the samples: for every expanded element its
            graph, its index in the compact (per-pair) array and its index inside its graph."""
            def __init__(lv, m_pair):
                lv.m_pair = cnt(m_pair)
                lv.cbase = excl(lv.m_pair)                              # compact base of every pair
                lv.m_graph = lv.m_pair[pair_of_graph]
                lv.gbase = excl(lv.m_graph)                             # expanded base of every graph [B + 1]
                lv.total = int(np.sum(m_pair)) * S
                lv.graph = torch.arange(B, **i64).repeat_interleave(lv.m_graph, output_size=lv.total)
                lv.local = torch.arange(lv.total, **i64) - lv.gbase[lv.graph]
                lv.src = lv.cbase[pair_of_graph[lv.graph]] + lv.local

        atoms, phs, bonds, rots, pps = Level(n_p), Level(P_p), Level(nb_p), Level(nrot_p), Level(npp_p)
        i32 = lambda t: t.to(torch.int32).contiguous()
        a0, p0 = atoms.gbase, phs.gbase                                # first atom / phore node of every graph
        self.lig_ptr, self.ph_ptr, self.rot_ptr = i32(a0), i32(p0), i32(rots.gbase)
        # first atom / rotatable bond of every group of TILE_GROUP graphs (+ total): restart points of dp_build_tiles
        G = self.tile_group = tile_group_for(int(np.sum(n_p)) * S, B)
        self.n_groups = (B + G - 1) // G
        self.lig_gptr = i32(torch.cat([a0[:-1:G], a0[-1:]]))
        self.rot_gptr = i32(torch.cat([rots.gbase[:-1:G], rots.gbase[-1:]]))
        self.lig_batch = i32(atoms.graph)
        # ---- bonds (CSR by source atom), rotatable bonds, phore-phore edges (CSR by source node)
        bptr_c = cat_ptr = up(np.concatenate([q.bond_ptr[:-1] for q in pa]).astype(np.int64))
        self.bond_ptr = i32(torch.cat([bptr_c[atoms.src] + bonds.gbase[atoms.graph], bonds.gbase[-1:]]))
        self.bond_dst = i32(cat('bond_dst', np.int64)[bonds.src] + a0[bonds.graph])
        self.bond_type = i32(cat('bond_type', np.int64)[bonds.src])
        self.rot_u = i32(cat('rot_u', np.int64)[rots.src] + a0[rots.graph])
        self.rot_v = i32(cat('rot_v', np.int64)[rots.src] + a0[rots.graph])
        self.pp_src = i32(cat('pp_src', np.int64)[pps.src] + p0[pps.graph])
        self.pp_dst = i32(cat('pp_dst', np.int64)[pps.src] + p0[pps.graph])
        pptr_c = up(np.concatenate([q.pp_ptr[:-1] for q in pa]).astype(np.int64))
        self.pp_ptr = i32(torch.cat([pptr_c[phs.src] + pps.gbase[phs.graph], pps.gbase[-1:]]))
        # ---- complete bipartite cross edges, sorted by (lig, phore) (smp:770-781) and the phore-major transposed order
        n_g, P_g = atoms.m_graph, phs.m_graph
        cbase = excl(n_g * P_g)
        self.cross_ptr = i32(cbase)
        cg = torch.arange(B, **i64).repeat_interleave(n_g * P_g, output_size=self.n_cross)
        cl_ = torch.arange(self.n_cross, **i64) - cbase[cg]
        ng, Pg = n_g[cg], P_g[cg]
        self.cross_lig, self.cross_ph = i32(a0[cg] + cl_ // Pg), i32(p0[cg] + cl_ % Pg)
        at, qt = cl_ % ng, cl_ // ng
        self.cross_lig_t, self.cross_ph_t = i32(a0[cg] + at), i32(p0[cg] + qt)
        self.cross_perm_t = i32(cbase[cg] + at * Pg + qt)
        # CSR of cross edges by ligand atom (canonical order) and by phore node (transposed order)
        self.cross_seg_lig = i32(excl(P_g[atoms.graph]))
        self.cross_seg_ph = i32(excl(n_g[phs.graph]))
        del cg, cl_, ng, Pg, at, qt

        # ---- node-aligned tiles of the static edge sets (dp_conv_fused); a node with > 256 edges: unfused kernels
        def tiles(per_pair, node_base, n_nodes):
            if any(t is None for t in per_pair):
                return None
            lv = Level(np.asarray([len(t) for t in per_pair]))
            tn = up(np.concatenate([np.asarray(t, np.int64) for t in per_pair]))[lv.src] + node_base[lv.graph]
            return (i32(torch.cat([tn, torch.full((1,), n_nodes, **i64)])), None, lv.total)

        def tiles_any(per_pair, node_base, n_nodes, deg_pair, nodes_pair):
            """Per-graph tiles expanded on the device when they are (nearly) full anyway; otherwise tiles that span the graphs
            of a group (grouped_tiles on the host, uploaded): sparse sets such as the phore-phore edges of an 8-point
            pharmacophore (24 edges per graph) fill 4x fewer M = 128 operands that way."""
            if any(t is None for t in per_pair):
                return None
            n_t = sum(len(t) for t in per_pair) * S
            n_e = int(sum(int(d.sum()) for d in deg_pair)) * S
            if n_t == 0 or n_e >= 0.95 * cap_edges * n_t:
                return tiles(per_pair, node_base, n_nodes)
            deg = np.concatenate([np.tile(np.asarray(d, np.int64), S) for d in deg_pair])
            tn = grouped_tiles(deg, np.repeat(nodes_pair, S), group=self.tile_group)
            return (i32(torch.cat([up(tn), torch.full((1,), n_nodes, **i64)])), None, len(tn))

        cap_edges = TILE_EDGES
        self.tiles_cross_lig = tiles_any(t_lig, a0, self.n_lig, [np.full(q.n, q.P) for q in pa], n_p)
        self.tiles_cross_ph = tiles_any(t_ph, p0, self.n_ph, [np.full(q.P, q.n) for q in pa], P_p)
        self.tiles_pp = tiles_any(t_pp, p0, self.n_ph, [np.diff(q.pp_ptr) for q in pa], P_p)
        # final_conv: one edge per atom to its graph's centre node -> output nodes = graphs, degree = atoms of the graph
        tn = grouped_tiles(np.repeat(n_p, S), np.ones(B, np.int64), group=64)
        self.tiles_final = None if tn is None else (i32(torch.cat([up(tn), torch.full((1,), B, **i64)])), None, len(tn))
        # ---- mask_rotate rows (uint8) and their per-graph byte offsets
        if on_host:
            # (the largest expansion of a small job - n_rot x n_atoms bytes per graph: the S copies of a pair's rows are S memcpys
            # here instead of three int64 index arrays of that size and a gather; same layout)
            self.mask, self.mask_off = expand_mask_rows_host(pa, S)
        else:
            msk = Level(nrot_p * n_p)
            self.mask = up(np.concatenate([q.mask.reshape(-1) for q in pa]).astype(np.uint8))[msk.src].contiguous() \
                if msk.total else torch.zeros(0, dtype=torch.uint8, device=device)
            self.mask_off = msk.gbase[:-1].contiguous()
        self.lig_arange = torch.arange(self.n_lig, dtype=torch.int32, device=device)
        # ---- node-level features
        f32c = lambda key: up(np.concatenate([getattr(q, key) for q in pa], 0).astype(np.float32))
        self.pos, self.norm = f32c('pos')[atoms.src].contiguous(), f32c('norm')[atoms.src].contiguous()
        self.phorefp, self.na1, self.na2 = (f32c(kk)[atoms.src].contiguous() for kk in ('phorefp', 'na1', 'na2'))
        self.ppos, self.pnorm, self.ptype = (f32c(kk)[phs.src].contiguous() for kk in ('ppos', 'pnorm', 'ptype'))
        # static parts of the AtomEncoders (setup-time gathers; smp:64-73), evaluated once per pair
        w = weights
        lig_tables, ph_tables, ph_lin_w = w.lig_tables, w.ph_tables, w.ph_lin_w
        if on_host:
            lig_tables = [w.h[f'encoder.lig_node_embedding.atom_embedding_list.{i}.weight'] for i in range(16)]
            ph_tables = [w.h[f'encoder.phore_node_embedding.atom_embedding_list.{i}.weight'] for i in range(3)]
            ph_lin_w = w.h['encoder.phore_node_embedding.linear.weight']
        x, px = up(np.concatenate([q.x for q in pa], 0).astype(np.int64)), f32c('px')
        ls = torch.zeros(x.shape[0], 20, device=device)
        for i in range(16):
            ls = ls + lig_tables[i][x[:, i]]
        ps = torch.zeros(px.shape[0], 20, device=device)
        for i in range(3):
            ps = ps + ph_tables[i][px[:, i].long()]
        # explicit products in a fixed order (a cuBLAS matmul may pick a size-dependent kernel => batch-composition-dependent rounding)
        ps = ps + px[:, 3:4] * ph_lin_w[:, 0][None, :] + px[:, 4:5] * ph_lin_w[:, 1][None, :]
        self.lig_static, self.ph_static = ls[atoms.src].contiguous(), ps[phs.src].contiguous()
        if on_host:
            # finished arrays -> device: laid out in ONE pinned staging buffer (256-byte aligned slots) and uploaded with ONE
            # asynchronous copy; the attributes become views of one device buffer.  (One pageable `.to()` per array - 65 of them -
            # cost ~7 ms of the ~10 ms a 40-graph job spends before its first kernel, tools/host_profile.py.)
            flat = []

            class _Ref:
                def __init__(self, i):
                    self.i = i

            def reg(v):
                if torch.is_tensor(v):
                    flat.append(v.contiguous())
                    return _Ref(len(flat) - 1)
                if isinstance(v, tuple):
                    return tuple(reg(t) for t in v)
                return v
            marked = {k: reg(v) for k, v in self.__dict__.items()}
            offs, tot = [], 0
            for t in flat:
                offs.append(tot)
                tot += (t.numel() * t.element_size() + 255) // 256 * 256
            stage = torch.empty(max(tot, 1), dtype=torch.uint8, pin_memory=final_device.type == 'cuda')
            for t, o in zip(flat, offs):
                nb = t.numel() * t.element_size()
                if nb:
                    stage[o:o + nb].copy_(t.reshape(-1).view(torch.uint8))
                    self.h2d_bytes += nb
            dbuf = torch.empty(max(tot, 1), dtype=torch.uint8, device=final_device)
            dbuf.copy_(stage, non_blocking=True)

            def res(v):
                if isinstance(v, _Ref):
                    t, o = flat[v.i], offs[v.i]
                    nb = t.numel() * t.element_size()
                    if nb == 0:
                        return torch.empty(t.shape, dtype=t.dtype, device=final_device)
                    return dbuf[o:o + nb].view(t.dtype).view(t.shape)
                if isinstance(v, tuple):
                    return tuple(res(t) for t in v)
                return v
            for k, v in marked.items():
                self.__dict__[k] = res(v)


class Workspace:
    """All per-step device buffers for one PackedBatch."""

    def __init__(self, b, weights, wbuf=None, fused=True):
        dev = b.device
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        i = lambda *s: torch.empty(*s, dtype=torch.int32, device=dev)
        self.lig_h = [f(b.n_lig, d) for d in LAYER_DIMS]
        self.ph_h = [f(b.n_ph, d) for d in LAYER_DIMS[:4]]
        self.thr, self.deg, self.gcount, self.gstart = i(b.n_lig), i(max(b.n_lig, b.n_rot, 1)), i(b.B), i(b.B + 1)
        self.ll_ptr, self.ll_src, self.ll_dst = i(b.n_lig + 1), i(b.ll_cap), i(b.ll_cap)
        self.ll_emb, self.ll_sh, self.ll_n = f(b.ll_cap, 20), f(b.ll_cap, 9), i(1)
        self.pp_h, self.pp_sh, self.pp_emb = f(b.n_pp, 20), f(b.n_pp, 9), f(b.n_pp, 20)
        self.cross_h, self.cross_fm, self.cross_tw = f(b.n_cross, 20), f(b.n_cross), f(b.n_cross)
        self.cross_emb, self.cross_sh, self.cross_nsh = f(b.n_cross, 20), f(b.n_cross, 9), f(b.n_cross, 9)
        self.c_emb, self.c_sh, self.gpred = f(b.n_lig, 20), f(b.n_lig, 9), f(b.B, 12)
        self.t_ptr, self.t_atom, self.t_u, self.t_v = i(b.n_rot + 1), i(max(b.tor_cap, 1)), i(max(b.tor_cap, 1)), i(max(b.tor_cap, 1))
        self.t_emb, self.t_sh, self.t_n = f(max(b.tor_cap, 1), 20), f(max(b.tor_cap, 1), 8), i(1)
        self.tor_feat = f(max(b.n_rot, 1), 40)
        self.tr, self.rot, self.tor = f(b.B, 3), f(b.B, 3), f(max(b.n_rot, 1))
        # HBM scratch of the unfused kernels (dp_edge_mlp(_tc) -> dp_tp_scatter): always needed by final_conv (hid = 40), by
        # every edge set when DIFFPHORE_CONV=split, and by static edge sets with a node of more than 128 edges
        split = [(b.n_lig, 200)]
        if not fused:
            split += [(b.ll_cap, 2200), (b.tor_cap, 1600)]
        if not fused or b.tiles_cross_lig is None or b.tiles_cross_ph is None:
            split.append((b.n_cross, 2200))
        if not fused or b.tiles_pp is None:
            split.append((b.n_pp, 1600))
        w_elems = max(max(e * w for e, w in split), 1)
        self.w_elems = w_elems
        max_edges = max(max(e for e, w in split if w != 200) if len(split) > 1 else 0, 1)
        self.hbuf = f(((max_edges + 127) // 128) * 128 * 64)          # hidden activations of dp_edge_mlp_tc (pass 1 -> pass 2)
        self.wbuf = wbuf if wbuf is not None and wbuf.numel() >= w_elems else f(w_elems)   # per-edge TP weights
        # node-aligned tiles of the dynamic edge sets (dp_build_tiles): <= 2 E / 128 + 1 tiles per graph
        self.tile_cnt, self.tile_start = i(b.B), i(b.B + 1)
        self.ll_tile_cap, self.tor_tile_cap = b.ll_cap // 64 + b.B, b.tor_cap // 64 + b.B
        self.ll_tiles, self.ll_ntiles = i(self.ll_tile_cap + 1), i(1)
        self.tor_tiles, self.tor_ntiles = i(self.tor_tile_cap + 1), i(1)
        self.n_launches = 0


class Engine:
    """Runs the score model and the conformer update for one PackedBatch on the current CUDA stream."""

    def __init__(self, weights):
        self.w = weights
        self.lib = weights.lib
        self.timer = None          # profiling.KernelTimer or None
        # second MLP layer on tcgen05 tensor cores (3xTF32) or on CUDA cores (FFMA); env DIFFPHORE_EDGE_MLP=ffma|tc
        import os
        self.use_tc = os.environ.get('DIFFPHORE_EDGE_MLP', 'tc') == 'tc'
        # DIFFPHORE_CONV=fused (default): dp_conv_fused, the per-edge hidden activations and weights never leave the SM;
        # DIFFPHORE_CONV=split: dp_edge_mlp(_tc) -> HBM -> dp_tp_scatter (kept for A/B measurements and as the path for
        # edge sets the fused kernel does not cover: hid != 60, W % 100 != 0, nodes with more than 128 edges)
        self.use_fused = os.environ.get('DIFFPHORE_CONV', 'fused') == 'fused'
        # see forward(): ligand / pharmacophore convolutions of a layer on two streams (the sampler switches it on for small chunks)
        self.concurrent = False
        self.side_stream = torch.cuda.Stream() if weights.device.type == 'cuda' else None

    def pack(self, graphs, samples_per_graph=1, wbuf=None):
        """Upload a batch and run the static-geometry setup kernels.  wbuf: optional shared per-edge weight buffer
        (chunks run one after the other on one stream, so they can share it)."""
        b = PackedBatch(graphs, samples_per_graph, self.w, self.w.device)
        ws = Workspace(b, self.w, wbuf, self.use_fused)
        st = torch.cuda.current_stream().cuda_stream
        sw = self.w.sw
        L.check(self.lib.dp_pp_setup(L.ptr(b.ppos), L.ptr(b.pp_src), L.ptr(b.pp_dst), b.n_pp, sw, L.ptr(ws.pp_h),
                                     L.ptr(ws.pp_sh), st), 'dp_pp_setup')
        L.check(self.lib.dp_cross_setup(L.ptr(b.cross_lig), L.ptr(b.cross_ph), b.n_cross, L.ptr(b.phorefp), L.ptr(b.ptype),
                                        sw, L.ptr(ws.cross_h), L.ptr(ws.cross_fm), st), 'dp_cross_setup')
        return b, ws

    # ------------------------------------------------------------------ one TensorProductConvLayer
    def _conv(self, cw, ws, emb, perm, tb, idxB, tc, idxC, idxC2, n_dev, n_cap, node_in, gather, sh, sh_stride, seg, out,
              residual, res_dim, mode, n_out, st, name='', tiles=None):
        p = L.ptr
        tm = self.timer
        if tm is not None:
            n_rec = n_cap
            if n_dev is not None:                       # dynamic edge count: stream-ordered copy to pinned memory
                n_rec = torch.empty(1, dtype=torch.int32).pin_memory()
                n_rec.copy_(n_dev, non_blocking=True)
            e0 = tm.start()
        if self.use_fused and tiles is not None and cw.w2img112 is not None:
            tile_node, n_tiles_dev, n_tiles_cap = tiles
            flat = getattr(cw, 'w2imgflat', None)                      # weight layout of the first-generation kernel (DIFFPHORE_W2)
            fn = self.lib.dp_conv_fused if flat is None else self.lib.dp_conv_fused_flat
            if getattr(cw, 'gen', 1) == 2:
                flat, fn = cw.w2img96, self.lib.dp_conv_fused2
                cw.flat_mode_bits = 0
            L.check(fn(cw.layer_id, p(emb), p(perm), p(tb), p(idxB), tb.shape[1], p(tc), p(idxC), p(idxC2),
                                           tc.shape[1], p(cw.w1img), cw.inv_w1scale, p(cw.w2img112 if flat is None else flat), cw.inv_wscale, p(node_in),
                                           p(gather), p(sh), sh_stride, p(seg), p(tile_node), p(n_tiles_dev), n_tiles_cap,
                                           p(cw.oscale), p(cw.oshift), p(out), p(residual), res_dim,
                                           mode if flat is None else mode | cw.flat_mode_bits, st), 'dp_conv_fused')
            if tm is not None:
                # columns the tensor pipe really multiplies per edge (profiling.roofline_fused: issued MMA work)
                if getattr(cw, 'gen', 1) == 2:
                    nch = -(-cw.W // 96)
                    mma_cols = (nch - 1) * 96 + -(-(cw.W - (nch - 1) * 96) // 16) * 16
                elif flat is None:
                    mma_cols = cw.W * 1.12
                else:
                    nch = -(-cw.W // 112)
                    mma_cols = (nch - 1) * 112 + (-(-(cw.W - (nch - 1) * 112) // 16) * 16 if cw.flat_mode_bits else 112)
                tm.stop('conv_fused', name, e0, n_rec, dict(W=cw.W, hid=cw.hid, in_dim=cw.in_dim, d_in=cw.d_in, d_out=cw.d_out, mma_cols=mma_cols,
                                                            n_out=n_out, tp_flops=cw.tp_flops))
            ws.n_launches += 1
            return
        if self.use_tc and cw.w2img is not None:
            L.check(self.lib.dp_edge_mlp_tc(p(emb), p(perm), p(tb), p(idxB), tb.shape[1], p(tc), p(idxC), p(idxC2), tc.shape[1],
                                            p(cw.w1), p(cw.b1), p(cw.w2img), cw.inv_wscale, cw.in_dim, cw.hid, cw.W, p(n_dev),
                                            n_cap, p(ws.hbuf), p(ws.wbuf), st), 'dp_edge_mlp_tc')
        else:
            L.check(self.lib.dp_edge_mlp(p(emb), p(perm), p(tb), p(idxB), tb.shape[1], p(tc), p(idxC), p(idxC2),
                                         tc.shape[1] if tc is not None else 0, p(cw.w1), p(cw.b1), p(cw.w2t), cw.in_dim,
                                         cw.hid, cw.W, p(n_dev), n_cap, p(ws.wbuf), st), 'dp_edge_mlp')
        if tm is not None:
            tm.stop('edge_mlp', name, e0, n_rec, dict(in_dim=cw.in_dim, hid=cw.hid, W=cw.W))
            e0 = tm.start()
        L.check(self.lib.dp_tp_scatter(cw.layer_id, p(node_in), p(gather), p(perm), p(sh), sh_stride, p(ws.wbuf), p(seg),
                                       p(cw.oscale), p(cw.oshift), p(out), p(residual), res_dim, mode, n_out, st),
                'dp_tp_scatter')
        if tm is not None:
            tm.stop('tp_scatter', name, e0, n_rec, dict(W=cw.W, d_in=cw.d_in, d_out=cw.d_out, d_sh=sh_stride if sh_stride == 9 else 7,
                                                        n_out=n_out))
        ws.n_launches += 2

    def forward(self, b, ws, sc):
        """Score model for the batch's current pose.  sc: device tensor [256] (ModelWeights.step_consts).
        Results land in ws.tr [B,3], ws.rot [B,3], ws.tor [n_rot]."""
        lib, sw, p = self.lib, self.w.sw, L.ptr
        st = torch.cuda.current_stream().cuda_stream
        scp = p(sc)
        L.check(lib.dp_node_embed(p(b.pos), p(b.ppos), p(b.lig_batch), p(b.ph_ptr), p(b.ptype), p(b.lig_static),
                                  p(b.ph_static), b.n_lig, b.n_ph, sw, scp, p(ws.lig_h[0]), p(ws.ph_h[0]), st), 'dp_node_embed')
        L.check(lib.dp_lig_graph(p(b.pos), p(b.lig_ptr), p(b.bond_ptr), p(b.bond_dst), p(b.bond_type), b.B, b.n_lig,
                                 b.max_atoms, sw, scp, p(ws.thr), p(ws.deg), p(ws.gcount), p(ws.gstart), p(ws.ll_ptr),
                                 p(ws.ll_src), p(ws.ll_dst), p(ws.ll_emb), p(ws.ll_sh), p(ws.ll_n), st), 'dp_lig_graph')
        L.check(lib.dp_pp_step(p(ws.pp_h), b.n_pp, sw, scp, p(ws.pp_emb), st), 'dp_pp_step')
        L.check(lib.dp_cross_step(p(b.pos), p(b.norm), p(b.ppos), p(b.pnorm), p(b.lig_ptr), p(b.ph_ptr), p(b.cross_ptr),
                                  b.B, b.max_atoms, p(b.phorefp), p(b.ptype), p(b.na1), p(b.na2), p(ws.cross_h),
                                  p(ws.cross_fm), sw, scp, p(ws.cross_tw), p(ws.cross_emb), p(ws.cross_sh),
                                  p(ws.cross_nsh), st), 'dp_cross_step')
        ws.n_launches += 6
        ll_tiles = None
        if self.use_fused:
            L.check(lib.dp_build_tiles(p(ws.ll_ptr), p(b.lig_gptr), b.n_groups, p(ws.tile_cnt), p(ws.tile_start), p(ws.ll_tiles),
                                       p(ws.ll_ntiles), st), 'dp_build_tiles')
            ws.n_launches += 3
            ll_tiles = (ws.ll_tiles, ws.ll_ntiles, ws.ll_tile_cap)
        cv = self.w.convs
        # Small (latency-bound) batches: the three convolutions that update the pharmacophore nodes of a layer are independent of the
        # three that update the ligand nodes (both read layer l, write different arrays) -> second stream, joined at the end of
        # the layer (captured as a fork / join inside the CUDA graphs).  Every output keeps its own order of additions.
        # The per-edge weight scratch of the unfused fallback is shared, so only the all-fused path may overlap.
        side = self.side_stream if (self.concurrent and self.timer is None and self.use_fused and b.tiles_cross_lig is not None
                                    and b.tiles_cross_ph is not None and b.tiles_pp is not None) else None
        main = torch.cuda.current_stream()
        for l in range(4):
            lh, ph, lo = ws.lig_h[l], ws.ph_h[l] if l < 4 else None, ws.lig_h[l + 1]
            d = LAYER_DIMS[l]
            st2 = st
            if l != 3 and side is not None:
                side.wait_stream(main)
                st2 = side.cuda_stream
            self._conv(cv[('lig', l)], ws, ws.ll_emb, None, lh, ws.ll_src, lh, ws.ll_dst, None, ws.ll_n, b.ll_cap,
                       lh, ws.ll_dst, ws.ll_sh, 9, ws.ll_ptr, lo, lh, d, 1, b.n_lig, st, f'lig{l}', ll_tiles)
            self._conv(cv[('phore_to_lig', l)], ws, ws.cross_emb, None, lh, b.cross_lig, ph, b.cross_ph, None, None,
                       b.n_cross, ph, b.cross_ph, ws.cross_sh, 9, b.cross_seg_lig, lo, None, 0, 2, b.n_lig, st, f'p2l{l}',
                       b.tiles_cross_lig)
            self._conv(cv[('phore_to_lig_norm', l)], ws, ws.cross_emb, None, lh, b.cross_lig, ph, b.cross_ph, None, None,
                       b.n_cross, ph, b.cross_ph, ws.cross_nsh, 9, b.cross_seg_lig, lo, None, 0, 2, b.n_lig, st, f'p2ln{l}',
                       b.tiles_cross_lig)
            if l != 3:
                po = ws.ph_h[l + 1]
                self._conv(cv[('phore', l)], ws, ws.pp_emb, None, ph, b.pp_src, ph, b.pp_dst, None, None, b.n_pp,
                           ph, b.pp_dst, ws.pp_sh, 9, b.pp_ptr, po, ph, d, 1, b.n_ph, st2, f'pp{l}', b.tiles_pp)
                self._conv(cv[('lig_to_phore', l)], ws, ws.cross_emb, b.cross_perm_t, lh, b.cross_lig_t, ph, b.cross_ph_t,
                           None, None, b.n_cross, lh, b.cross_lig_t, ws.cross_sh, 9, b.cross_seg_ph, po, None, 0, 2,
                           b.n_ph, st2, f'l2p{l}', b.tiles_cross_ph)
                self._conv(cv[('lig_to_phore_norm', l)], ws, ws.cross_emb, b.cross_perm_t, lh, b.cross_lig_t, ph,
                           b.cross_ph_t, None, None, b.n_cross, lh, b.cross_lig_t, ws.cross_nsh, 9, b.cross_seg_ph, po,
                           None, 0, 2, b.n_ph, st2, f'l2pn{l}', b.tiles_cross_ph)
                if side is not None:
                    main.wait_stream(side)
        h4 = ws.lig_h[4]
        L.check(lib.dp_center_step(p(b.pos), p(b.lig_ptr), b.B, sw, scp, p(ws.c_emb), p(ws.c_sh), st), 'dp_center_step')
        if self.use_fused and cv['final'].w2img112 is not None and b.tiles_final is not None:
            self._conv(cv['final'], ws, ws.c_emb, None, h4, b.lig_arange, h4, b.lig_arange, None, None, b.n_lig, h4, None, ws.c_sh, 9,
                       b.lig_ptr, ws.gpred, None, 0, 0, b.B, st, 'final', b.tiles_final)
        else:
            self._conv(cv['final'], ws, ws.c_emb, None, h4, b.lig_arange, None, None, None, None, b.n_lig, h4, None, ws.c_sh, 9,
                       b.lig_ptr, ws.gpred, None, 0, 0, b.B, st, 'final')
        L.check(lib.dp_score_head(p(ws.gpred), b.B, sw, scp, p(ws.tr), p(ws.rot), st), 'dp_score_head')
        ws.n_launches += 2
        if b.n_rot > 0:
            L.check(lib.dp_tor_graph(p(b.pos), p(b.lig_ptr), p(b.rot_ptr), p(b.rot_u), p(b.rot_v), b.B, b.n_rot, sw,
                                     p(ws.deg), p(ws.gcount), p(ws.gstart), p(ws.t_ptr), p(ws.t_atom), p(ws.t_u), p(ws.t_v),
                                     p(ws.t_emb), p(ws.t_sh), p(ws.t_n), st), 'dp_tor_graph')
            tor_tiles = None
            if self.use_fused:
                L.check(lib.dp_build_tiles(p(ws.t_ptr), p(b.rot_gptr), b.n_groups, p(ws.tile_cnt), p(ws.tile_start), p(ws.tor_tiles),
                                           p(ws.tor_ntiles), st), 'dp_build_tiles')
                ws.n_launches += 3
                tor_tiles = (ws.tor_tiles, ws.tor_ntiles, ws.tor_tile_cap)
            self._conv(cv['tor'], ws, ws.t_emb, None, h4, ws.t_atom, h4, ws.t_u, ws.t_v, ws.t_n, b.tor_cap, h4, ws.t_atom,
                       ws.t_sh, 8, ws.t_ptr, ws.tor_feat, None, 0, 0, b.n_rot, st, 'tor', tor_tiles)
            L.check(lib.dp_tor_head(p(ws.tor_feat), b.n_rot, sw, scp, p(ws.tor), st), 'dp_tor_head')
            ws.n_launches += 4
        return ws.tr, ws.rot, ws.tor[:b.n_rot]

    def update(self, b, ws, sc, tr_z=None, rot_z=None, tor_z=None, no_torsion=False):
        """Apply the Euler–Maruyama step (sampling.py:223-254) to b.pos / b.norm in place."""
        p = L.ptr
        st = torch.cuda.current_stream().cuda_stream
        L.check(self.lib.dp_conformer_update(p(b.pos), p(b.norm), p(b.lig_ptr), p(b.rot_ptr), p(b.rot_u), p(b.rot_v),
                                             p(b.mask), p(b.mask_off), b.B, b.max_atoms, b.max_rot, p(ws.tr), p(ws.rot),
                                             p(ws.tor), p(tr_z), p(rot_z), p(tor_z), p(sc), int(no_torsion), st),
                'dp_conformer_update')
        ws.n_launches += 1

    def randomize(self, b, tor_init, rot_init, tr_init, no_torsion=False):
        p = L.ptr
        st = torch.cuda.current_stream().cuda_stream
        L.check(self.lib.dp_randomize_position(p(b.pos), p(b.norm), p(b.lig_ptr), p(b.rot_ptr), p(b.rot_u), p(b.rot_v),
                                               p(b.mask), p(b.mask_off), b.B, b.max_atoms, b.max_rot, p(tor_init),
                                               p(rot_init), p(tr_init), int(no_torsion), st), 'dp_randomize_position')
