"""ORACLE (test infrastructure, not product code) — parity pinned only indirectly, see DESIGN.md §Oracle.

CPU (PyTorch fp32/fp64) restatement of DiffPhore's score network for the shipped flag set
(weights/diffphore_calibrated_warmuped_ft/model_parameters.yml), following
/root/reference/src/models/score_model_phore.py line by line:

    TensorProductScoreModel.forward            smp:294-310
    get_trtheta_score                          smp:313-378
    build_center_conv_graph                    smp:381-406
    build_bond_conv_graph                      smp:409-437
    LigPhoreEncoder.forward                    smp:644-712
    build_lig_conv_graph / build_phore_conv_graph   smp:715-756
    _build_phoretype_cross_conv_graph          smp:759-895
    boarder_analyze                            smp:898-935
    AtomEncoder / TensorProductConvLayer / GaussianSmearing / angle_vectors /
    fully_connect_two_graphs / my_sort_edge_index   smp:23-149, 978-1097

Third-party operators (e3nn / torch_cluster / torch_scatter / PyG) are restated in e3nn_lite.py and below
(radius graph: brute force, squared distance < r^2, at most 32 neighbours per centre keeping the LOWEST
indices (torch_cluster's CUDA rule)).

The model reads its parameters straight from a reference-format state_dict (385 tensors), so the shipped
checkpoint is consumed unchanged.  `data` is any PyG-HeteroDataBatch-like object (duck typed).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import e3nn_lite as e3

LIG_FEATURE_DIMS = [119, 4, 12, 12, 8, 10, 6, 6, 2, 8, 2, 2, 2, 2, 2, 2]   # process_mols.py:162-179
PHORE_FEATURE_DIMS = [11, 2, 2]                                             # process_pharmacophore.py:34-38


def default_config():
    """The subset of model_parameters.yml that get_model (src/utils/utils.py:113-168) forwards."""
    return dict(ns=20, nv=10, num_conv_layers=4, sigma_embed_dim=20, distance_embed_dim=20,
                cross_distance_embed_dim=20, lig_max_radius=5.0, cross_max_distance=25.0,
                center_max_distance=30.0, embedding_scale=10000, scaler=100.0,
                clash_cutoff=[1.0, 2.0, 3.0, 4.0, 5.0], num_phoretype=11, max_neighbors=32,
                tr_sigma_min=0.1, tr_sigma_max=5.0, rot_sigma_min=0.1, rot_sigma_max=1.5,
                tor_sigma_min=0.0314, tor_sigma_max=3.14, no_clamp=False)


def irrep_seq(ns, nv):
    return [f'{ns}x0e', f'{ns}x0e + {nv}x1o', f'{ns}x0e + {nv}x1o + {nv}x1e',
            f'{ns}x0e + {nv}x1o + {nv}x1e + {ns}x0o']                        # smp:586-591


def sinusoidal_embedding(timesteps, embedding_dim, max_positions=10000):
    """src/utils/diffusion_utils.py:82-93 (the frequency table is always built in fp32 there)."""
    half = embedding_dim // 2
    emb = math.log(max_positions) / (half - 1)
    emb = torch.exp(torch.arange(half, dtype=torch.float32) * -emb)
    emb = timesteps.float()[:, None] * emb[None, :]
    return torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)


def gaussian_smearing(dist, start, stop, n):
    """smp:978-1015: offset=linspace(start,stop,n) (fp32), coeff=-0.5/(offset[1]-offset[0])**2."""
    offset = torch.linspace(start, stop, n)
    coeff = -0.5 / (offset[1] - offset[0]).item() ** 2
    d = dist.view(-1, 1) - offset.to(dist.dtype).view(1, -1)
    return torch.exp(coeff * d.pow(2))


def radius_pairs(x, y, r, batch_x, batch_y, max_num_neighbors=32):
    """torch_cluster.radius(x, y, r, batch_x, batch_y): for every query y_i the points x_j of the same graph with
    |x_j - y_i|^2 < r^2, lowest j first, at most `max_num_neighbors`.  Returns [2,E] (row0 -> y, row1 -> x)."""
    d2 = ((y[:, None, :] - x[None, :, :]) ** 2).sum(-1)
    ok = (d2 < r * r) & (batch_y[:, None] == batch_x[None, :])
    rank = torch.cumsum(ok.long(), dim=1)
    ok &= rank <= max_num_neighbors
    yi, xj = torch.nonzero(ok, as_tuple=True)
    return torch.stack([yi, xj], 0)


def radius_graph(pos, r, batch, max_num_neighbors=32):
    """torch_cluster.radius_graph(flow='source_to_target', loop=False): radius(x, x, ..., max+1) INCLUDING the
    self pair, which is removed afterwards; returned as [neighbour j (source), centre i (target)]."""
    e = radius_pairs(pos, pos, r, batch, batch, max_num_neighbors + 1)
    e = e[:, e[0] != e[1]]
    return torch.stack([e[1], e[0]], 0)


def scatter(src, index, dim_size, reduce='sum'):
    out = src.new_zeros((dim_size,) + src.shape[1:])
    out.index_add_(0, index, src)
    if reduce == 'mean':
        cnt = torch.zeros(dim_size, dtype=src.dtype).index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
        out = out / cnt.clamp(min=1).view(-1, *([1] * (src.dim() - 1)))
    return out


def angle_vectors(a, b):
    a_norm = a.norm(dim=-1, keepdim=True)
    b_norm = b.norm(dim=-1, keepdim=True)
    return 2 * torch.atan2((a * b_norm - a_norm * b).norm(dim=-1), (a * b_norm + a_norm * b).norm(dim=-1))


class OracleScoreModel:
    """forward(data) -> (tr_pred [B,3], rot_pred [B,3], tor_pred [sum n_rot]) like smp:294-310.

    so3_norm / torus_norm: callables sigma(np.ndarray) -> np.ndarray standing for
    utils.so3.score_norm / utils.torus.score_norm (smp:352,376) so that the (unseeded Monte-Carlo, H1)
    tables can be injected identically into the oracle and the CUDA path."""

    def __init__(self, state_dict, so3_norm, torus_norm, config=None, dtype=torch.float32):
        self.cfg = default_config()
        if config:
            self.cfg.update(config)
        self.dtype = dtype
        self.sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in state_dict.items()}
        self.so3_norm, self.torus_norm = so3_norm, torus_norm
        ns, nv = self.cfg['ns'], self.cfg['nv']
        self.seq = [e3.parse_irreps(s) for s in irrep_seq(ns, nv)]
        self.sh = e3.sh_irreps(2)
        self.trace = None      # set to a dict to capture intermediates

    # ------------------------------------------------------------------ small helpers
    def t_to_sigma(self, t_tr, t_rot, t_tor):
        c = self.cfg                                                         # diffusion_utils.py:16-20
        return (c['tr_sigma_min'] ** (1 - t_tr) * c['tr_sigma_max'] ** t_tr,
                c['rot_sigma_min'] ** (1 - t_rot) * c['rot_sigma_max'] ** t_rot,
                c['tor_sigma_min'] ** (1 - t_tor) * c['tor_sigma_max'] ** t_tor)

    def temb(self, t):
        return sinusoidal_embedding(self.cfg['embedding_scale'] * t, self.cfg['sigma_embed_dim']).to(self.dtype)

    def linear(self, prefix, x):
        y = x @ self.sd[prefix + '.weight'].T
        if prefix + '.bias' in self.sd:
            y = y + self.sd[prefix + '.bias']
        return y

    def mlp(self, prefix, x, act=F.relu, last=3, final_act=None):
        """nn.Sequential(Linear, act, Dropout, Linear[, final_act]) -> modules 0 and 3."""
        y = self.linear(f'{prefix}.{last}', act(self.linear(f'{prefix}.0', x)))
        return final_act(y) if final_act is not None else y

    def atom_encoder(self, prefix, x, n_cat):
        emb = 0
        for i in range(n_cat):                                                # smp:64-73
            emb = emb + self.sd[f'{prefix}.atom_embedding_list.{i}.weight'][x[:, i].long()]
        if x.shape[1] > n_cat:
            emb = emb + self.linear(f'{prefix}.linear', x[:, n_cat:].to(self.dtype))
        return emb

    def _rec(self, name, val):
        if self.trace is not None:
            self.trace[name] = val.detach().clone()

    # ------------------------------------------------------------------ TensorProductConvLayer (smp:134-149)
    def conv(self, prefix, in_irreps, sh_irreps_, out_irreps, node_attr, edge_index, edge_attr, edge_sh, out_nodes):
        instrs, numel = e3.fctp_instructions(in_irreps, sh_irreps_, out_irreps)
        assert numel == self.sd[prefix + '.fc.3.weight'].shape[0], (prefix, numel)
        edge_src, edge_dst = edge_index
        w = self.mlp(prefix + '.fc', edge_attr)
        tp = e3.fctp_apply(in_irreps, sh_irreps_, out_irreps, instrs, node_attr[edge_dst], edge_sh, w)
        out = scatter(tp, edge_src, out_nodes, 'mean')
        self._rec(prefix + '.pre_bn', out)
        out = e3.batchnorm_eval(out, out_irreps, self.sd[prefix + '.batch_norm.weight'],
                                self.sd[prefix + '.batch_norm.bias'], self.sd[prefix + '.batch_norm.running_mean'],
                                self.sd[prefix + '.batch_norm.running_var'])
        self._rec(prefix + '.out', out)
        return out

    # ------------------------------------------------------------------ graph builders
    def build_lig_conv_graph(self, data):
        c, lig = self.cfg, data['ligand']
        lig.node_sigma_emb = self.temb(lig.node_t['tr'])                      # smp:717
        pos = lig.pos.to(self.dtype)
        radius_edges = radius_graph(pos, c['lig_max_radius'], lig.batch, c['max_neighbors'])
        bond_index = data['ligand', 'ligand'].edge_index.long()
        edge_index = torch.cat([bond_index, radius_edges], 1)
        edge_attr = torch.cat([data['ligand', 'ligand'].edge_attr.to(self.dtype),
                               torch.zeros(radius_edges.shape[1], 4, dtype=self.dtype)], 0)
        edge_attr = torch.cat([edge_attr, lig.node_sigma_emb[edge_index[0]]], 1)
        node_attr = torch.cat([lig.x.to(self.dtype), lig.node_sigma_emb], 1)
        src, dst = edge_index
        edge_vec = pos[dst] - pos[src]
        edge_attr = torch.cat([edge_attr, gaussian_smearing(edge_vec.norm(dim=-1), 0.0, c['lig_max_radius'],
                                                            c['distance_embed_dim'])], 1)
        return node_attr, edge_index, edge_attr, e3.spherical_harmonics(edge_vec)

    def build_phore_conv_graph(self, data):
        c, ph = self.cfg, data['phore']
        ph.node_sigma_emb = self.temb(ph.node_t['tr'])                        # smp:744
        node_attr = torch.cat([ph.x.to(self.dtype), ph.node_sigma_emb], 1)
        edge_index = data['phore', 'phore'].edge_index.long()
        src, dst = edge_index
        pos = ph.pos.to(self.dtype)
        edge_vec = pos[dst] - pos[src]
        edge_attr = torch.cat([ph.node_sigma_emb[src],
                               gaussian_smearing(edge_vec.norm(dim=-1), 0.0, 5.0, c['distance_embed_dim'])], 1)
        return node_attr, edge_index, edge_attr, e3.spherical_harmonics(edge_vec)

    def cross_edges(self, data):
        """fully_connect_two_graphs (non-EX then EX) + my_sort_edge_index == all (lig, phore) pairs of a graph
        sorted by (lig idx, phore idx) (smp:770-781, 1038-1097)."""
        lb, pb = data['ligand'].batch, data['phore'].batch
        same = lb[:, None] == pb[None, :]
        src, dst = torch.nonzero(same, as_tuple=True)
        return torch.stack([src, dst], 0)

    def build_cross_conv_graph(self, data):
        c, lig, ph = self.cfg, data['ligand'], data['phore']
        T = c['num_phoretype']
        edge_index = self.cross_edges(data)
        src, dst = edge_index
        lpos, ppos = lig.pos.to(self.dtype), ph.pos.to(self.dtype)
        phoretype, phorefp = ph.phoretype.to(self.dtype), lig.phorefp.to(self.dtype)
        pnorm = ph.norm.to(self.dtype)
        edge_vec = ppos[dst] - lpos[src]
        edge_length_emb = gaussian_smearing(edge_vec.norm(dim=-1), 0.0, c['cross_max_distance'],
                                            c['cross_distance_embed_dim'])
        edge_attr = torch.cat([lig.node_sigma_emb[src], edge_length_emb], 1)
        is_ex = phoretype[dst, -1] == 1
        aggreement = phoretype[dst] * phorefp[src]                           # smp:790-793
        aggreement = torch.where(is_ex[:, None], torch.zeros_like(aggreement), aggreement)
        phoretype_attr = torch.cat([aggreement, phoretype[dst], phorefp[src]], -1)
        # new=True, phoretype_match=True, all three *_transition=True  (smp:799-848)
        distance = self.mlp('encoder.cross_distance_transition', edge_length_emb, final_act=F.softplus)
        feat_match = self.mlp('encoder.phoretype_match_transition', phoretype_attr, final_act=F.softplus)
        total_weight = feat_match * distance * c['scaler']
        dirv = self.mlp('encoder.phore_direction_transition', total_weight, act=F.leaky_relu, final_act=F.leaky_relu)
        direction = torch.pow(-1.0, (dirv < 0).to(self.dtype))
        edge_vec = edge_vec * direction
        ex = total_weight.exp()                                               # atom_weight == 'phore' (smp:835-840)
        atom_weight = ex / scatter(ex, src, lig.pos.shape[0], 'sum')[src]
        total_weight = atom_weight                                            # multiple=False (H3, smp:845)
        edge_vec = edge_vec * total_weight
        edge_attr = torch.cat([edge_attr, phoretype_attr], -1)                # use_phore_match_feat (smp:860-862)
        # angle_match (smp:874-889)
        lnorm_all = lig.norm.to(self.dtype)[src].reshape(-1, T, 3)
        lig_norm = torch.sum(aggreement.unsqueeze(-1) * lnorm_all, dim=1)
        cr = torch.cross(lig_norm, pnorm[dst], dim=-1)        # H6: identical to dim-less cross unless E_c == 3
        if not c['no_clamp']:
            cr = torch.clip(cr, 1e-12)                                        # H10: component-wise clamp
        rotate_norm = F.normalize(cr * torch.sum(aggreement, dim=-1, keepdim=True))
        curr_angle = angle_vectors(lig_norm, pnorm[dst]).unsqueeze(-1)
        a1 = torch.sum(aggreement * lig.norm_angle1.to(self.dtype)[src], dim=-1, keepdim=True)
        a2 = torch.sum(aggreement * lig.norm_angle2.to(self.dtype)[src], dim=-1, keepdim=True)
        d1, d2 = curr_angle - a1, curr_angle - a2
        # torch.sort(...).indices[:,0] of [|d1|,|d2|]: index 0 unless |d2| < |d1| (stable on ties)
        norm_real = torch.where(d2.abs() < d1.abs(), d2, d1)
        if getattr(self, 'branch_log', None) is not None:
            # H7 diagnostics (tests only): the discrete decisions of this block per cross edge and their margins.  `active`:
            # edges whose rotate_norm is not identically zero (some phore type agrees with the atom's fingerprint).
            raw = torch.cross(lig_norm, pnorm[dst], dim=-1)
            self.branch_log.append(dict(src=src.clone(), active=aggreement.sum(-1) != 0, clamped=raw < 1e-12,
                                        choice=(d2.abs() < d1.abs()).squeeze(-1),
                                        clamp_margin=(raw - 1e-12).abs(),
                                        choice_margin=torch.where(a1 == a2, torch.full_like(d1, float('inf')),
                                                                  (d2.abs() - d1.abs()).abs()).squeeze(-1)))
        rotate_norm = rotate_norm * norm_real
        self._rec('cross.edge_vec', edge_vec)
        self._rec('cross.rotate_norm', rotate_norm)
        return edge_index, edge_attr, e3.spherical_harmonics(edge_vec), e3.spherical_harmonics(rotate_norm)

    def boarder_analyze(self, data):
        """smp:898-935: per atom the distance to the nearest exclusion sphere of its graph (+1e9 if none) and the
        five `<= cutoff` flags.  (The reference goes through to_dense_batch + torch.cdist; direct differences
        are used here, see DESIGN.md hazard H11.)"""
        lig, ph = data['ligand'], data['phore']
        lpos, ppos = lig.pos.to(self.dtype), ph.pos.to(self.dtype)
        d = (lpos[:, None, :] - ppos[None, :, :]).norm(dim=-1)
        ok = (lig.batch[:, None] == ph.batch[None, :]) & (ph.phoretype[:, -1] == 1)[None, :]
        d = d + (1 - ok.to(self.dtype)) * 1e9
        dis_min = d.min(dim=-1).values.unsqueeze(-1)
        clashed = dis_min.tile([1, len(self.cfg['clash_cutoff'])]) <= torch.tensor(self.cfg['clash_cutoff'], dtype=self.dtype)
        return torch.cat([clashed.to(self.dtype), dis_min], -1)

    # ------------------------------------------------------------------ encoder (smp:644-712)
    def encoder(self, data):
        ns = self.cfg['ns']
        lig_node_attr, lig_edge_index, lig_edge_attr, lig_edge_sh = self.build_lig_conv_graph(data)
        lig_src, lig_dst = lig_edge_index
        lig_node_attr = self.atom_encoder('encoder.lig_node_embedding', lig_node_attr, len(LIG_FEATURE_DIMS))
        lig_node_attr = lig_node_attr + self.atom_encoder('encoder.boarder_embedding', self.boarder_analyze(data), 5)
        lig_edge_attr = self.mlp('encoder.lig_edge_embedding', lig_edge_attr)

        phore_node_attr, phore_edge_index, phore_edge_attr, phore_edge_sh = self.build_phore_conv_graph(data)
        phore_src, phore_dst = phore_edge_index
        phore_node_attr = self.atom_encoder('encoder.phore_node_embedding', phore_node_attr, len(PHORE_FEATURE_DIMS))
        phore_edge_attr = self.mlp('encoder.phore_edge_embedding', phore_edge_attr)

        cross_edge_index, cross_edge_attr, cross_edge_sh, cross_edge_norm_sh = self.build_cross_conv_graph(data)
        cross_lig, cross_phore = cross_edge_index
        cross_edge_attr = self.mlp('encoder.cross_edge_embedding', cross_edge_attr)
        cross_flip = torch.flip(cross_edge_index, dims=[0])
        self._rec('lig_node_attr0', lig_node_attr); self._rec('phore_node_attr0', phore_node_attr)
        self._rec('lig_edge_index', lig_edge_index); self._rec('lig_edge_attr', lig_edge_attr)
        self._rec('lig_edge_sh', lig_edge_sh); self._rec('phore_edge_attr', phore_edge_attr)
        self._rec('cross_edge_attr', cross_edge_attr); self._rec('cross_edge_sh', cross_edge_sh)
        self._rec('cross_edge_norm_sh', cross_edge_norm_sh)

        L = self.cfg['num_conv_layers']
        nL, nP = lig_node_attr.shape[0], phore_node_attr.shape[0]
        for l in range(L):
            ii, oi = self.seq[min(l, 3)], self.seq[min(l + 1, 3)]
            lig_edge_attr_ = torch.cat([lig_edge_attr, lig_node_attr[lig_src, :ns], lig_node_attr[lig_dst, :ns]], -1)
            lig_intra = self.conv(f'encoder.lig_conv_layers.{l}', ii, self.sh, oi, lig_node_attr, lig_edge_index,
                                  lig_edge_attr_, lig_edge_sh, nL)
            p2l_attr = torch.cat([cross_edge_attr, lig_node_attr[cross_lig, :ns], phore_node_attr[cross_phore, :ns]], -1)
            lig_inter = self.conv(f'encoder.phore_to_lig_conv_layers.{l}', ii, self.sh, oi, phore_node_attr,
                                  cross_edge_index, p2l_attr, cross_edge_sh, nL)
            lig_inter_norm = self.conv(f'encoder.phore_to_lig_norm_conv_layers.{l}', ii, self.sh, oi, phore_node_attr,
                                       cross_edge_index, p2l_attr, cross_edge_norm_sh, nL)
            if l != L - 1:
                phore_edge_attr_ = torch.cat([phore_edge_attr, phore_node_attr[phore_src, :ns],
                                              phore_node_attr[phore_dst, :ns]], -1)
                phore_intra = self.conv(f'encoder.phore_conv_layers.{l}', ii, self.sh, oi, phore_node_attr,
                                        phore_edge_index, phore_edge_attr_, phore_edge_sh, nP)
                phore_inter = self.conv(f'encoder.lig_to_phore_conv_layers.{l}', ii, self.sh, oi, lig_node_attr,
                                        cross_flip, p2l_attr, cross_edge_sh, nP)
                phore_inter_norm = self.conv(f'encoder.lig_to_phore_norm_conv_layers.{l}', ii, self.sh, oi,
                                             lig_node_attr, cross_flip, p2l_attr, cross_edge_norm_sh, nP)
            lig_node_attr = F.pad(lig_node_attr, (0, lig_intra.shape[-1] - lig_node_attr.shape[-1]))
            lig_node_attr = lig_node_attr + lig_intra + lig_inter + lig_inter_norm
            if l != L - 1:
                phore_node_attr = F.pad(phore_node_attr, (0, phore_intra.shape[-1] - phore_node_attr.shape[-1]))
                phore_node_attr = phore_node_attr + phore_intra + phore_inter + phore_inter_norm
            self._rec(f'lig_node_attr{l + 1}', lig_node_attr)
        return lig_node_attr, phore_node_attr

    # ------------------------------------------------------------------ score heads (smp:313-378)
    def forward(self, data):
        c, ns = self.cfg, self.cfg['ns']
        lig = data['ligand']
        lig_node_attr, _ = self.encoder(data)
        tr_sigma, rot_sigma, tor_sigma = self.t_to_sigma(*[data.complex_t[k] for k in ('tr', 'rot', 'tor')])
        B = data.num_graphs
        pos = lig.pos.to(self.dtype)
        # build_center_conv_graph (smp:381-406)
        n_atoms = lig.batch.shape[0]
        edge_index = torch.stack([lig.batch.long(), torch.arange(n_atoms)], 0)
        center_pos = torch.zeros(B, 3, dtype=self.dtype).index_add_(0, lig.batch.long(), pos)
        center_pos = center_pos / torch.bincount(lig.batch.long(), minlength=B).unsqueeze(1)
        edge_vec = pos[edge_index[1]] - center_pos[edge_index[0]]
        edge_attr = gaussian_smearing(edge_vec.norm(dim=-1), 0.0, c['center_max_distance'], c['distance_embed_dim'])
        edge_attr = torch.cat([edge_attr, lig.node_sigma_emb[edge_index[1]]], 1)
        edge_sh = e3.spherical_harmonics(edge_vec)
        edge_attr = self.mlp('center_edge_embedding', edge_attr)
        edge_attr = torch.cat([edge_attr, lig_node_attr[edge_index[1], :ns]], -1)
        global_pred = self.conv('final_conv', self.seq[3], self.sh, e3.parse_irreps('2x1o + 2x1e'), lig_node_attr,
                                edge_index, edge_attr, edge_sh, B)
        tr_pred = global_pred[:, :3] + global_pred[:, 6:9]
        rot_pred = global_pred[:, 3:6] + global_pred[:, 9:]
        graph_sigma_emb = self.temb(data.complex_t['tr'])
        tr_norm = torch.linalg.vector_norm(tr_pred, dim=1).unsqueeze(1)
        tr_pred = tr_pred / tr_norm * self._final_mlp('tr_final_layer', torch.cat([tr_norm, graph_sigma_emb], 1))
        rot_norm = torch.linalg.vector_norm(rot_pred, dim=1).unsqueeze(1)
        rot_pred = rot_pred / rot_norm * self._final_mlp('rot_final_layer', torch.cat([rot_norm, graph_sigma_emb], 1))
        tr_pred = tr_pred / tr_sigma.to(self.dtype).unsqueeze(1)
        so3n = torch.from_numpy(np.asarray(self.so3_norm(rot_sigma.float().numpy()))).float()   # smp:352 (.float())
        rot_pred = rot_pred * so3n.to(self.dtype).unsqueeze(1)
        edge_mask = lig.edge_mask.bool()
        if edge_mask.sum() == 0:
            return tr_pred, rot_pred, torch.empty(0, dtype=self.dtype)

        # build_bond_conv_graph (smp:409-437)
        bond_index = data['ligand', 'ligand'].edge_index.long()
        bonds = bond_index[:, edge_mask]
        bond_pos = (pos[bonds[0]] + pos[bonds[1]]) / 2
        bond_batch = lig.batch[bonds[0]]
        tor_edge_index = radius_pairs(pos, bond_pos, c['lig_max_radius'], lig.batch, bond_batch, c['max_neighbors'])
        tvec = pos[tor_edge_index[1]] - bond_pos[tor_edge_index[0]]
        tor_edge_attr = gaussian_smearing(tvec.norm(dim=-1), 0.0, c['lig_max_radius'], c['distance_embed_dim'])
        tor_edge_attr = self.mlp('final_edge_embedding', tor_edge_attr)
        tor_edge_sh = e3.spherical_harmonics(tvec)
        tor_bond_vec = pos[bonds[1]] - pos[bonds[0]]
        tor_bond_attr = lig_node_attr[bonds[0]] + lig_node_attr[bonds[1]]
        tor_bonds_sh = e3.spherical_harmonics(tor_bond_vec, only_l=2)
        sh45_irreps, tor_edge_sh = e3.full_tp_apply(self.sh, [(1, 2, 1)], tor_edge_sh, tor_bonds_sh[tor_edge_index[0]])
        tor_edge_attr = torch.cat([tor_edge_attr, lig_node_attr[tor_edge_index[1], :ns],
                                   tor_bond_attr[tor_edge_index[0], :ns]], -1)
        n_rot = int(edge_mask.sum())
        tor_pred = self.conv('tor_bond_conv', self.seq[3], sh45_irreps, e3.parse_irreps(f'{ns}x0o + {ns}x0e'),
                             lig_node_attr, tor_edge_index, tor_edge_attr, tor_edge_sh, n_rot)
        tor_pred = (torch.tanh(tor_pred @ self.sd['tor_final_layer.0.weight'].T)
                    @ self.sd['tor_final_layer.3.weight'].T).squeeze(1)
        edge_sigma = tor_sigma[lig.batch.long()][bond_index[0]][edge_mask]
        tn = torch.tensor(np.asarray(self.torus_norm(edge_sigma.float().numpy()))).float()     # smp:376 (.float())
        tor_pred = tor_pred * torch.sqrt(tn).to(self.dtype)
        return tr_pred, rot_pred, tor_pred

    def _final_mlp(self, prefix, x):
        """nn.Sequential(Linear, Dropout, ReLU, Linear) -> modules 0 and 3 (smp:265-266)."""
        return self.linear(prefix + '.3', F.relu(self.linear(prefix + '.0', x)))

    __call__ = forward
