"""GPU-box aid: run-to-run and chunking determinism of the sampler (bit-exact expected)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'src'))
from diffphore_b200.engine import ModelWeights
from diffphore_b200.sampler import DenoisingSampler
from diffphore_b200.synthetic import make_pairs
from tests.parity_util import random_state_dict, make_draws
P, steps = 17, int(sys.argv[1]) if len(sys.argv) > 1 else 3
graphs = make_pairs(P, 32, 8)
sd = random_state_dict(0)
dev = torch.device('cuda:0')
init, noise, n_rot = make_draws(graphs, 1, 3, steps=steps)
w = ModelWeights(sd, dev)
def run(**kw):
    smp = DenoisingSampler(w, steps, **kw)
    tr = []
    pos, ptr = smp.run(graphs, 1, noise=noise, init=init, trace=tr)
    return pos, tr, len(smp.prepare(graphs, 1))
a, ta, na = run()
b, tb, nb = run()
c, tc_, nc = run(resident_bytes=4 << 20, weight_buffer_bytes=int(os.environ.get("WB", 64 << 20)))
print('chunks', na, nb, nc)
print('run-to-run   : equal', torch.equal(a, b), 'max diff', float((a - b).abs().max()))
print('chunked vs 1 : equal', torch.equal(a, c), 'max diff', float((a - c).abs().max()))
n0 = tc_[0][0].shape[0]
nr0 = tc_[0][2].shape[0]
for k in range(steps):                      # chunked trace order: chunk 0 step 0..steps-1, chunk 1 ...
    d = [float((x[:n] - y).abs().max()) for x, y, n in zip(ta[k], tc_[k], (n0, n0, nr0))]
    bad = (ta[k][0][:n0] != tc_[k][0]).any(1).nonzero().flatten().tolist()
    print(f'step {k}: first-chunk score diffs tr/rot/tor', d, 'graphs with different tr', bad)

ptr = np.concatenate([[0], np.cumsum([g['ligand'].pos.shape[0] for g in graphs])])
badg = [i for i in range(P) if not torch.equal(a[ptr[i]:ptr[i + 1]], c[ptr[i]:ptr[i + 1]])]
print('graphs with different final poses:', badg)
rot_off = np.concatenate([[0], np.cumsum(n_rot)])
for ch in range(1, nc):
    g0, g1 = 8 * ch, min(8 * ch + 8, P)
    for k in range(steps):
        t = tc_[ch * steps + k]
        d = [float((ta[k][0][g0:g1] - t[0]).abs().max()), float((ta[k][1][g0:g1] - t[1]).abs().max()),
             float((ta[k][2][rot_off[g0]:rot_off[g1]] - t[2]).abs().max()) if t[2].numel() else 0.0]
        print(f'chunk {ch} step {k}: score diffs tr/rot/tor', d)
print('n_rot per graph', n_rot)
