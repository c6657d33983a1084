"""GPU-box aid: dp_conv_fused against dp_edge_mlp_tc + dp_tp_scatter on one synthetic convolution
(same inputs, same weights), with CUDA-event timings of both pipelines."""
import math, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'src'))
from diffphore_b200 import lib as L
from diffphore_b200.engine import _make_w2img, _make_w2img112, _make_w2imgflat, _make_w1img, greedy_tiles

FLAT = os.environ.get('DIFFPHORE_W2', 'paths') == 'flat'      # time the flat weight layout (dp_conv_fused_flat)
GEN2 = os.environ.get('DIFFPHORE_CONV_GEN', '1') == '2'       # time dp_conv_fused2 (96-column flat chunks)
NARROW = os.environ.get('PROBE_WINDOW', 'narrow') == 'narrow'  # gather sources within 60 consecutive rows per pair tile (like one or two graphs)
import numpy as np
lib = L.load(); p = L.ptr
dev = torch.device('cuda:0')
CFG = {0: (20, 50, 600, 9), 1: (50, 80, 1100, 9), 2: (80, 100, 1600, 9), 3: (100, 100, 2200, 9), 5: (100, 40, 1600, 8)}


def run(layer, n_nodes, deg, seed=0, mode=0, reps=3):
    d_in, d_out, W, shs = CFG[layer]
    g = torch.Generator().manual_seed(seed)
    degs = np.full(n_nodes, deg) if isinstance(deg, int) else np.asarray(deg)
    seg = np.concatenate([[0], np.cumsum(degs)]).astype(np.int32)
    E = int(seg[-1])
    n_in = 5000
    emb = torch.randn(E, 20, generator=g).to(dev)
    nodes = torch.randn(n_in, d_in, generator=g).to(dev)
    nodes20 = torch.randn(n_in, 100, generator=g).to(dev)
    ib = torch.randint(0, n_in, (E,), generator=g, dtype=torch.int32).to(dev)
    ic = torch.randint(0, n_in, (E,), generator=g, dtype=torch.int32).to(dev)
    gat = torch.randint(0, n_in, (E,), generator=g, dtype=torch.int32)
    if NARROW:
        gat = ((torch.arange(E) // 256 * 7) % (n_in - 100) + torch.randint(0, 60, (E,), generator=g)).to(torch.int32)
    gat = gat.to(dev)
    sh = torch.randn(E, shs, generator=g).to(dev)
    w1c, b1c = torch.randn(60, 60, generator=g) / 8, torch.randn(60, generator=g)
    w1, b1 = w1c.to(dev), b1c.to(dev)
    img1, inv1 = _make_w1img(w1c, b1c); img1 = img1.to(dev)
    w3, b3 = torch.randn(W, 60, generator=g) / 8, torch.randn(W, generator=g)
    img, inv_ws = _make_w2img(w3, b3); img = img.to(dev)
    img112, inv2 = (_make_w2imgflat if FLAT else _make_w2img112)(w3, b3)
    if GEN2:
        img112, inv2 = _make_w2imgflat(w3, b3, 96)
    img112 = img112.to(dev)
    assert inv_ws == inv2
    oscale, oshift = torch.rand(d_out, generator=g).to(dev) + 0.5, torch.randn(d_out, generator=g).to(dev)
    segd = torch.from_numpy(seg).to(dev)
    tiles = greedy_tiles(degs)
    tile_node = torch.tensor(tiles + [n_nodes], dtype=torch.int32, device=dev)
    n_tiles = len(tiles)
    hbuf = torch.empty(((E + 127) // 128) * 128 * 64, device=dev)
    wbuf = torch.empty(E * W, device=dev)
    res = torch.randn(n_nodes, d_in, generator=g).to(dev)
    out0 = torch.randn(n_nodes, d_out, generator=g).to(dev)
    st = torch.cuda.current_stream().cuda_stream

    def split(out):
        L.check(lib.dp_edge_mlp_tc(p(emb), None, p(nodes20), p(ib), 100, p(nodes20), p(ic), None, 100, p(w1), p(b1), p(img), inv_ws,
                                   60, 60, W, None, E, p(hbuf), p(wbuf), st), 'mlp')
        L.check(lib.dp_tp_scatter(layer, p(nodes), p(gat), None, p(sh), shs, p(wbuf), p(segd), p(oscale), p(oshift), p(out),
                                  p(res), d_in, mode, n_nodes, st), 'tp')

    def fused(out):
        L.check((lib.dp_conv_fused2 if GEN2 else (lib.dp_conv_fused_flat if FLAT else lib.dp_conv_fused))(layer, p(emb), None, p(nodes20), p(ib), 100, p(nodes20), p(ic), None, 100, p(img1), inv1, p(img112),
                                  inv_ws, p(nodes), p(gat), p(sh), shs, p(segd), p(tile_node), None, n_tiles, p(oscale), p(oshift),
                                  p(out), p(res), d_in, mode, st), 'fused')

    oa, ob = out0.clone(), out0.clone()
    split(oa); fused(ob)
    torch.cuda.synchronize()
    err = float((oa - ob).norm() / oa.norm())
    mx = float((oa - ob).abs().max())
    if err > 1e-5:
        bad = ((oa - ob).abs().max(1).values > 1e-3 * oa.abs().max()).nonzero().flatten().cpu().numpy()
        tl = np.asarray(tiles + [n_nodes])
        print('   BAD nodes', len(bad), 'of', n_nodes, 'first', bad[:12], 'deg', degs[bad[:12]], 'tile', np.searchsorted(tl, bad[:12], 'right') - 1,
              'seg', seg[bad[:12]], 'n_tiles', n_tiles)
    tm = {}
    for name, fn in (('split', split), ('fused', fused)):
        o = out0.clone()
        for _ in range(2):
            fn(o)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn(o)
        e1.record(); torch.cuda.synchronize()
        tm[name] = e0.elapsed_time(e1) / reps
    fl = E * (2.0 * 61 * W)
    print(f'layer {layer} mode {mode} E={E} nodes={n_nodes} tiles={n_tiles}: rel-L2 {err:.2e} max-abs {mx:.2e} | split {tm["split"]:.3f} ms '
          f'fused {tm["fused"]:.3f} ms ({fl / tm["fused"] / 1e9:.0f} TFLOP/s alg, {E / tm["fused"] / 1e3:.1f} Medges/s)', flush=True)
    return err


if __name__ == '__main__' and '--stamps' not in sys.argv:
    small = '--small' in sys.argv
    errs = []
    errs.append(run(0, 64, 8, mode=0))
    errs.append(run(0, 37, [3, 0, 17, 128, 1] * 7 + [5, 9], mode=2))
    if not small:
        for layer in (0, 1, 2, 3, 5):
            errs.append(run(layer, 148 * 16 * 16, 8, mode=1 if layer != 5 else 0))
        rng = np.random.default_rng(0)
        errs.append(run(3, 20000, rng.integers(0, 40, 20000), mode=2))
    print('max rel err', max(errs))
    assert max(errs) < 1e-5


def stamps(layer=3):
    """Per-item clock stamps of CTA 0's second tile pair: MMA issuer vs the two drain groups."""
    import ctypes
    raw = ctypes.CDLL(L.LIB_PATH)
    dbg = torch.zeros(3 * 64 * 3, dtype=torch.int64, device=dev)
    raw.dp_debug_set_cf_probe(ctypes.c_void_p(dbg.data_ptr()))
    run(layer, 148 * 16 * 16, 8, mode=1 if layer != 5 else 0, reps=1)
    raw.dp_debug_set_cf_probe(ctypes.c_void_p(0))
    d = dbg.cpu().reshape(3, 64, 3)
    t0 = int(d[0, 0, 0])
    ph = [int(x) - t0 for x in d[1, 50]] + [int(x) - t0 for x in d[1, 51]]
    print('phases (warp 0): prologue start %d, A ready %d, main loop end %d, staged %d, reduced %d, pair end %d' % tuple(ph))
    print('epilogue (warp 0): after barrier 1 %d, rows staged %d, after barrier 2 %d' % tuple(int(x) - t0 for x in d[1, 52]))
    for i in range(22):
        m = [int(x) - t0 for x in d[0, i]]
        w0 = [int(x) - t0 for x in d[1, i]]
        w1 = [int(x) - t0 for x in d[2, i]]
        print(f'chunk {i:2d}: issuer0 wait {m[0]:6d}->{m[1]:6d} issued {m[2]:6d} | tile0 wait_full {w0[0]:6d}->{w0[1]:6d} drained {w0[2]:6d} | '
              f'tile1 wait_full {w1[0]:6d}->{w1[1]:6d} drained {w1[2]:6d}')


if __name__ == '__main__' and '--stamps' in sys.argv:
    stamps()
