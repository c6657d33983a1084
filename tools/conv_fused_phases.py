"""GPU-box aid: per-phase clock stamps of dp_conv_fused on the REAL cfg2 workload (last layer-3 convolution of one forward =
phore->lig norm conv), CTA 0, second tile pair."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'src'))
from diffphore_b200 import lib as L
from diffphore_b200.engine import ModelWeights, Engine
from diffphore_b200.synthetic import make_pairs
from diffphore_b200.tables import So3ScoreNorm, TorusScoreNorm
from tests.parity_util import random_state_dict
dev = torch.device('cuda:0')
w = ModelWeights(random_state_dict(0), dev)
eng = Engine(w)
b, ws = eng.pack(make_pairs(int(sys.argv[1]) if len(sys.argv) > 1 else 64, 32, 8), 40)
sc = w.step_consts(0.5, So3ScoreNorm(), TorusScoreNorm()).to(dev)
eng.forward(b, ws, sc); torch.cuda.synchronize()
raw = ctypes.CDLL(L.LIB_PATH)
dbg = torch.zeros(3 * 64 * 3, dtype=torch.int64, device=dev)
raw.dp_debug_set_cf_probe(ctypes.c_void_p(dbg.data_ptr()))
eng.forward(b, ws, sc); torch.cuda.synchronize()
raw.dp_debug_set_cf_probe(ctypes.c_void_p(0))
d = dbg.cpu().reshape(3, 64, 3)
t0 = int(d[0, 0, 0])
ph = [int(x) - t0 for x in d[1, 50]] + [int(x) - t0 for x in d[1, 51]] + [int(x) - t0 for x in d[1, 52]]
print('prologue start %d | A1 ready %d | main loop end %d | after load_attr %d | reduced %d | pair end %d | barrier1 %d | rows staged %d | barrier2 %d' % tuple(ph))
for i in (0, 1, 2, 10, 20, 21):
    m = [int(x) - t0 for x in d[0, i]]; w0 = [int(x) - t0 for x in d[1, i]]; w1 = [int(x) - t0 for x in d[2, i]]
    print(f'chunk {i:2d}: issuer0 wait {m[0]:6d}->{m[1]:6d} issued {m[2]:6d} | tile0 wait_full {w0[0]:6d}->{w0[1]:6d} drained {w0[2]:6d} | tile1 {w1[0]:6d}->{w1[1]:6d} drained {w1[2]:6d}')
