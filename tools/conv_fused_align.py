"""GPU-box aid: does dp_conv_fused give bit-identical node outputs when the tile alignment of the same edges changes?"""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conv_fused_probe import *
from diffphore_b200.engine import _make_w1img, _make_w2img112, greedy_tiles

def conv(layer, degs, shift, seed=0):
    """Random conv on nodes with degrees `degs`; `shift` extra leading nodes (deg 50) move every tile boundary."""
    d_in, d_out, W, shs = CFG[layer]
    g = torch.Generator().manual_seed(seed)
    n = len(degs); E = int(np.sum(degs))
    emb = torch.randn(E, 20, generator=g); nodes = torch.randn(5000, d_in, generator=g); nodes20 = torch.randn(5000, 100, generator=g)
    ib = torch.randint(0, 5000, (E,), generator=g, dtype=torch.int32); ic = torch.randint(0, 5000, (E,), generator=g, dtype=torch.int32)
    gat = torch.randint(0, 5000, (E,), generator=g, dtype=torch.int32); sh = torch.randn(E, shs, generator=g)
    w1c, b1c = torch.randn(60, 60, generator=g) / 8, torch.randn(60, generator=g)
    w3, b3 = torch.randn(W, 60, generator=g) / 8, torch.randn(W, generator=g)
    oscale, oshift = torch.rand(d_out, generator=g) + 0.5, torch.randn(d_out, generator=g)
    # prepend `shift` dummy nodes with 50 edges each (copies of the first edges)
    ex = 50 * shift
    cat = lambda t: torch.cat([t[:ex], t]) if ex else t
    emb, ib, ic, gat, sh = cat(emb), cat(ib), cat(ic), cat(gat), cat(sh)
    degs2 = np.concatenate([np.full(shift, 50, dtype=np.int64), np.asarray(degs)])
    seg = np.concatenate([[0], np.cumsum(degs2)]).astype(np.int32)
    tiles = greedy_tiles(degs2)
    img1, inv1 = _make_w1img(w1c, b1c); img112, inv2 = _make_w2img112(w3, b3)
    D = lambda t: t.to(dev)
    emb, nodes, nodes20, ib, ic, gat, sh, oscale, oshift, img1, img112 = map(D, (emb, nodes, nodes20, ib, ic, gat, sh, oscale, oshift, img1, img112))
    segd = torch.from_numpy(seg).to(dev); tile_node = torch.tensor(tiles + [len(degs2)], dtype=torch.int32, device=dev)
    out = torch.zeros(len(degs2), d_out, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    L.check(lib.dp_conv_fused(layer, p(emb), None, p(nodes20), p(ib), 100, p(nodes20), p(ic), None, 100, p(img1), inv1, p(img112),
                              inv2, p(nodes), p(gat), p(sh), shs, p(segd), p(tile_node), None, len(tiles), p(oscale), p(oshift),
                              p(out), None, 0, 0, st), 'fused')
    torch.cuda.synchronize()
    return out[shift:].cpu()

rng = np.random.default_rng(1)
for layer in (0, 3, 5):
    degs = rng.integers(1, 40, 3000)
    ref = conv(layer, degs, 0)
    for shift in (1, 2, 3):
        o = conv(layer, degs, shift)
        bad = (o != ref).any(1).nonzero().flatten()
        print(f'layer {layer} shift {shift}: bit-identical {torch.equal(o, ref)}  differing nodes {len(bad)}  max diff {float((o - ref).abs().max()):.3e}', bad[:8].tolist())
