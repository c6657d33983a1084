"""GPU-box profiling aid: per-phase clock64() stamps of one CTA of dp_edge_mlp_tc + a plain HBM write-bandwidth probe."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'src'))
from diffphore_b200 import lib as L
from diffphore_b200.engine import _make_w2img
lib = L.load(); p = L.ptr
raw = ctypes.CDLL(L.LIB_PATH)
dev = torch.device('cuda:0')
for W in (600, 2200):
    E = 128 * 148 * 8
    g = torch.Generator().manual_seed(0)
    emb = torch.randn(E, 20, generator=g).to(dev); nodes = torch.randn(5000, 100, generator=g).to(dev)
    ib = torch.randint(0, 5000, (E,), generator=g, dtype=torch.int32).to(dev); ic = torch.randint(0, 5000, (E,), generator=g, dtype=torch.int32).to(dev)
    w1, b1 = torch.randn(60, 60).to(dev), torch.randn(60).to(dev)
    img, inv_ws = _make_w2img(torch.randn(W, 60), torch.randn(W)); img = img.to(dev)
    out = torch.empty(E, W, device=dev)
    hbuf = torch.empty(E * 64, device=dev)
    dbg = torch.zeros(64, dtype=torch.int64, device=dev)
    raw.dp_debug_set_tc_probe(ctypes.c_void_p(dbg.data_ptr()))
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        L.check(lib.dp_edge_mlp_tc(p(emb), None, p(nodes), p(ib), 100, p(nodes), p(ic), None, 100, p(w1), p(b1), p(img), inv_ws, 60, 60, W, None, E, p(hbuf), p(out), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(lib.dp_edge_mlp_tc(p(emb), None, p(nodes), p(ib), 100, p(nodes), p(ic), None, 100, p(w1), p(b1), p(img), inv_ws, 60, 60, W, None, E, p(hbuf), p(out), st))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    d = dbg.cpu().tolist()
    for half in (0, 1, 2):
        t = d[half * 16:half * 16 + 16]
        print(f'W={W} half={half} stamps (cycles since start):', [x - d[0] if x else None for x in t])
    print(f'W={W}: {ms:.3f} ms, {E * W * 4 / ms / 1e6:.0f} GB/s write, tiles/SM={E // 128 // 148}')
    raw.dp_debug_set_tc_probe(ctypes.c_void_p(0))
    V = ctypes.c_void_p
    e0.record()
    raw.dp_debug_edge_hidden(V(emb.data_ptr()), V(nodes.data_ptr()), V(ib.data_ptr()), 100, V(nodes.data_ptr()), V(ic.data_ptr()), 100,
                             V(w1.data_ptr()), V(b1.data_ptr()), E, V(hbuf.data_ptr()), V(st))
    e1.record(); torch.cuda.synchronize()
    print(f'   pass 1 (hidden) alone: {e0.elapsed_time(e1):.3f} ms for {E} edges')
buf = torch.empty(2 << 30, dtype=torch.float32, device=dev)
for _ in range(2):
    buf.fill_(1.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); buf.fill_(2.0); e1.record(); torch.cuda.synchronize()
print(f'fill_ 8 GiB: {buf.numel() * 4 / e0.elapsed_time(e1) / 1e6:.0f} GB/s write-only')
src = torch.empty_like(buf)
e0.record(); buf.copy_(src); e1.record(); torch.cuda.synchronize()
print(f'copy_ 8 GiB: {2 * buf.numel() * 4 / e0.elapsed_time(e1) / 1e6:.0f} GB/s read+write')
