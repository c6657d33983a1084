"""Shared parity harness: CUDA path (through the C ABI) vs the CPU oracle on identical seeded inputs.

Used by tests/ (-m gpu), __graft_entry__.smoke() and tools/gpu_debug.py.  Tolerances (fp32 path):
  forward : rel-L2(tr), rel-L2(rot), rel-L2(tor) <= 1e-4 against the fp32 oracle
  update  : max |pos - pos_oracle| <= 2e-5 A x max(1, ligand extent / 8 A) for one conformer update (fp32 SVD of the reference)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'src')):
    if p not in sys.path:
        sys.path.insert(0, p)

LOCAL_CKPT = os.path.join(ROOT, 'oracle', '_ref', 'weights', 'best_ema_inference_epoch_model.pt')
SHIPPED_KW = dict(sigma_embed_dim=20, ns=20, nv=10, num_conv_layers=4, distance_embed_dim=20, cross_distance_embed_dim=20,
                  consider_norm=True, boarder=True, use_phore_match_feat=True, cross_distance_transition=True,
                  phore_direction_transition=True, phoretype_match_transition=True, atom_weight='phore',
                  auto_phorefp=False, scaler=100.0, dropout=0.1, clash_cutoff=[1.0, 2.0, 3.0, 4.0, 5.0])


def have_checkpoint():
    return os.path.exists(LOCAL_CKPT)


def real_state_dict():
    return torch.load(LOCAL_CKPT, map_location='cpu', weights_only=False)


from diffphore_b200.synthetic import random_state_dict  # noqa: E402,F401  (lives in the package: bench.py uses it too)


def load_pairs(kind, n_pairs, n_atoms=32, n_phore=8):
    from diffphore_b200.synthetic import make_pairs
    from diffphore_b200.graph import graph_from_arrays
    if kind == 'synthetic':
        return make_pairs(n_pairs, n_atoms, n_phore)
    a = np.load(os.path.join(ROOT, 'tests', 'golden', 'real_pairs.npz'))
    names = list(a['names'])
    return [graph_from_arrays(a, f'p{k}_', str(names[k])) for k in range(min(n_pairs, len(names)))]


def make_draws(graphs, samples, seed, tr_sigma_max=5.0, steps=0):
    """Seeded initial-pose draws (sampling.py:35,49,58) and per-step Gaussian noise (sampling.py:230-244)."""
    from scipy.spatial.transform import Rotation as R
    rng = np.random.RandomState(seed)
    n_rot = [int(g['ligand'].edge_mask.sum()) for g in graphs for _ in range(samples)]
    B = len(n_rot)
    init = dict(tor=rng.uniform(-np.pi, np.pi, size=sum(n_rot)).astype(np.float32),
                rot=R.random(B, random_state=rng).as_matrix().astype(np.float32),
                tr=(rng.randn(B, 3) * tr_sigma_max).astype(np.float32))
    noise = [dict(tr=rng.randn(B, 3).astype(np.float32), rot=rng.randn(B, 3).astype(np.float32),
                  tor=rng.randn(sum(n_rot)).astype(np.float32)) for _ in range(steps)]
    return init, noise, n_rot


def oracle_initial_graphs(graphs, samples, init, n_rot):
    """data_list of the reference (pair-major copies) after oracle randomize_position with the injected draws."""
    from oracle import sampler as osamp
    dl = [g.clone() for g in graphs for _ in range(samples)]
    offs = np.concatenate([[0], np.cumsum(n_rot)])
    osamp.randomize_position(dl, False, False, 5.0, [init['tor'][offs[i]:offs[i + 1]] for i in range(len(dl))],
                             list(init['rot']), list(init['tr']))
    return dl


def rel(a, b):
    a, b = torch.as_tensor(a).double().reshape(-1), torch.as_tensor(b).double().reshape(-1)
    if b.numel() == 0:
        return 0.0
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def run_forward_parity(n_pairs=2, n_atoms=12, n_phore=5, samples=2, weights='random', kind='synthetic', t=0.6, seed=0,
                       check_update=False, detail=False, device='cuda:0', tol=1e-4):
    from diffphore_b200.engine import ModelWeights, Engine
    from diffphore_b200.graph import collate
    from diffphore_b200.tables import So3ScoreNorm, TorusScoreNorm
    from oracle.model import OracleScoreModel
    from oracle import sampler as osamp
    sd = real_state_dict() if weights == 'real' else random_state_dict(seed)
    graphs = load_pairs(kind, n_pairs, n_atoms, n_phore)
    init, noise, n_rot = make_draws(graphs, samples, seed, steps=1)
    # draw poses at noise level t rather than t=1 so that the geometry is in-distribution for level t
    tr_s = 0.1 ** (1 - t) * 5.0 ** t
    init['tr'] = init['tr'] / 5.0 * tr_s
    dl = oracle_initial_graphs(graphs, samples, init, n_rot)
    so3n, torn = So3ScoreNorm(), TorusScoreNorm()
    om = OracleScoreModel(sd, so3n, torn)
    om.trace = {} if detail else None
    batch = collate([g.clone() for g in dl])
    osamp.set_time(batch, t, len(dl))
    o_tr, o_rot, o_tor = om(batch)

    dev = torch.device(device)
    w = ModelWeights(sd, dev)
    eng = Engine(w)
    b, ws = eng.pack(dl, 1)
    dt = 0.05
    sc = w.step_consts(t, so3n, torn, dt=dt).to(dev)
    tr, rot, tor = eng.forward(b, ws, sc)
    torch.cuda.synchronize()
    res = dict(tr=rel(tr.cpu(), o_tr), rot=rel(rot.cpu(), o_rot), tor=rel(tor.cpu(), o_tor), B=len(dl),
               n_lig=b.n_lig, n_rot=b.n_rot, launches=ws.n_launches)
    if detail:
        tr_ = om.trace
        res['h0'] = rel(ws.lig_h[0].cpu(), tr_['lig_node_attr0'])
        res['ph0'] = rel(ws.ph_h[0].cpu(), tr_['phore_node_attr0'])
        res['cross_emb'] = rel(ws.cross_emb.cpu(), tr_['cross_edge_attr'])
        res['cross_sh'] = rel(ws.cross_sh.cpu(), tr_['cross_edge_sh'])
        res['cross_nsh'] = rel(ws.cross_nsh.cpu(), tr_['cross_edge_norm_sh'])
        res['pp_emb'] = rel(ws.pp_emb.cpu(), tr_['phore_edge_attr'])
        n_e = int(ws.ll_n.cpu()[0])
        res['ll_edges'] = (n_e, tr_['lig_edge_index'].shape[1])
        # order-independent check of the ligand edge features: aggregate per aggregation node
        o_src = tr_['lig_edge_index'][0]
        agg_o = torch.zeros(b.n_lig, 29).index_add_(0, o_src, torch.cat([tr_['lig_edge_attr'], tr_['lig_edge_sh']], 1))
        m_src = ws.ll_src[:n_e].cpu().long()
        agg_m = torch.zeros(b.n_lig, 29).index_add_(0, m_src, torch.cat([ws.ll_emb[:n_e].cpu(), ws.ll_sh[:n_e].cpu()], 1))
        res['ll_feat'] = rel(agg_m, agg_o)
        for l in range(1, 5):
            res[f'lig_h{l}'] = rel(ws.lig_h[l].cpu(), tr_[f'lig_node_attr{l}'])
        res['gpred'] = rel(ws.gpred.cpu(), tr_['final_conv.out'])
        if b.n_rot:
            res['tor_feat'] = rel(ws.tor_feat[:b.n_rot].cpu(), tr_['tor_bond_conv.out'])
    ok = all(res[k] <= tol for k in ('tr', 'rot', 'tor'))
    if check_update:
        z = noise[0]
        zt = {k: torch.from_numpy(v).to(dev) for k, v in z.items()}
        eng.update(b, ws, sc, zt['tr'], zt['rot'], zt['tor'])
        torch.cuda.synchronize()
        # oracle update driven by the ORACLE scores
        c = om.cfg
        tr_g = (0.1 ** (1 - t) * 5.0 ** t) * np.sqrt(2 * np.log(c['tr_sigma_max'] / c['tr_sigma_min']))
        rot_g = 2 * (0.1 ** (1 - t) * 1.5 ** t) * np.sqrt(np.log(c['rot_sigma_max'] / c['rot_sigma_min']))
        tor_g = (0.0314 ** (1 - t) * 3.14 ** t) * np.sqrt(2 * np.log(c['tor_sigma_max'] / c['tor_sigma_min']))
        # use the CUDA scores for both sides so that this isolates the update kernel
        trp = (tr_g ** 2 * dt * tr.cpu() + tr_g * np.sqrt(dt) * torch.from_numpy(z['tr'])).float()
        rotp = (rot.cpu() * dt * rot_g ** 2 + rot_g * np.sqrt(dt) * torch.from_numpy(z['rot'])).float()
        torp = (tor_g ** 2 * dt * tor.cpu() + tor_g * np.sqrt(dt) * torch.from_numpy(z['tor'])).float().numpy()
        offs = np.concatenate([[0], np.cumsum(n_rot)])
        new = [osamp.modify_conformer(g.clone(), trp[i:i + 1], rotp[i], torp[offs[i]:offs[i + 1]]) for i, g in enumerate(dl)]
        o_pos = torch.cat([g['ligand'].pos for g in new], 0)
        o_norm = torch.cat([g['ligand'].norm for g in new], 0)
        res['upd_pos'] = float((b.pos.cpu() - o_pos).abs().max())
        res['upd_norm'] = float((b.norm.cpu() - o_norm).abs().max())
        # 2e-5 A for ligands up to 8 A across, growing with the lever arm beyond: the reference aligns the flexible onto the rigid
        # pose with a 3x3 SVD in FP32 (geometry.py:121-131; the kernel's Jacobi runs in fp64), whose rotation is good to a few
        # 1e-7 rad - times the distance of an atom from the centroid (15-25 A for the 64- and 128-atom test chains)
        cen = torch.cat([g['ligand'].pos.mean(0, keepdim=True).expand_as(g['ligand'].pos) for g in new], 0)
        res['extent'] = float((o_pos - cen).norm(dim=1).max())
        res['upd_tol'] = 2e-5 * max(1.0, res['extent'] / 8.0)
        ok = ok and res['upd_pos'] <= res['upd_tol'] and res['upd_norm'] <= 2e-5
    res['ok'] = bool(ok)
    return res
