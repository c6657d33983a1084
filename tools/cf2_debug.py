"""GPU-box aid: dp_conv_fused2 against dp_conv_fused (bit identity expected), with a report of where they differ."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'src'))
from diffphore_b200 import lib as L
from tests.test_gpu import _conv_case, _run_conv_fused, _CF


def case(layer, n_nodes, deg, window, mode, seed=0):
    rng = np.random.default_rng(seed)
    degs = np.full(n_nodes, deg) if isinstance(deg, int) else rng.integers(deg[0], deg[1], n_nodes)
    t = _conv_case(layer, degs, seed=seed)
    E = t['gat'].shape[0]
    if window == 'narrow':
        base = (torch.arange(E) // 256 * 7) % 600
        t['gat'] = (base + torch.randint(0, 60, (E,), generator=torch.Generator().manual_seed(seed))).to(torch.int32)
    d_in, d_out = _CF[layer][0], _CF[layer][1]
    g = torch.Generator().manual_seed(7)
    res, out0 = torch.randn(n_nodes, d_in, generator=g), torch.randn(n_nodes, d_out, generator=g)
    kw = {0: {}, 1: dict(residual=res), 2: dict(out0=out0)}[mode]
    ref = _run_conv_fused(layer, t, L, mode=mode, **kw)
    got = _run_conv_fused(layer, t, L, mode=mode, gen2=True, **kw)
    from diffphore_b200.engine import greedy_tiles
    tiles = np.asarray(greedy_tiles(t['degs']) + [n_nodes])
    bad = ((got - ref) != 0).any(1).nonzero().flatten().numpy()
    nan = int(torch.isnan(got).any(1).sum())
    print(f'layer {layer} nodes {n_nodes} deg {deg} {window} mode {mode}: tiles {len(tiles) - 1}, differing nodes {len(bad)} (nan rows {nan}), '
          f'max abs diff {float((got - ref).abs().nan_to_num(1e9).max()):.3e}, rel {float((got - ref).norm() / ref.norm()):.2e}', flush=True)
    if len(bad):
        tl = np.searchsorted(tiles, bad, 'right') - 1
        u, c = np.unique(tl, return_counts=True)
        print('   tiles with differences (tile: nodes; pair index = tile // 148):', [(int(a), int(b), int(a) // 148) for a, b in zip(u[:12], c[:12])], '... n =', len(u))
        cols = ((got - ref) != 0)[bad[0]].nonzero().flatten().numpy()
        print('   first bad node', bad[0], 'cols', cols[:20], 'n cols', len(cols), 'got', got[bad[0], cols[:4]].numpy(), 'ref', ref[bad[0], cols[:4]].numpy())
    return len(bad)


if __name__ == '__main__':
    L.load()
    layers = [int(x) for x in os.environ.get('LAYERS', '0,3').split(',')]
    for layer in layers:
        case(layer, 32 * 100, 8, 'narrow', 0)          # 100 tiles: one pair per CTA
        case(layer, 32 * 148 * 2, 8, 'narrow', 0)      # two pairs per CTA
        case(layer, 32 * 148 * 2, 8, 'wide', 0)
        case(layer, 32 * 148 * 3, 8, 'narrow', 1)
        case(layer, 32 * 148 * 4, 8, 'narrow', 2)
        case(layer, 32 * 148 * 8, 8, 'narrow', 0)
        case(layer, 6000, (0, 40), 'narrow', 0)
