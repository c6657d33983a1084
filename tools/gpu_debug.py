"""GPU-box debugging aid: stage-by-stage parity of the CUDA path against the oracle, written to gpurun_out/."""
import json, os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.parity_util import run_forward_parity, have_checkpoint

os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
cases = [dict(n_pairs=2, n_atoms=12, n_phore=5, samples=2, weights='random', t=0.6, check_update=True),
         dict(n_pairs=3, n_atoms=32, n_phore=8, samples=2, weights='random', t=0.3, check_update=True)]
if have_checkpoint():
    cases += [dict(n_pairs=4, kind='real', samples=2, weights='real', t=0.5, check_update=True),
              dict(n_pairs=2, n_atoms=32, n_phore=8, samples=2, weights='real', t=0.8)]
out = []
for c in cases:
    t0 = time.time()
    try:
        r = run_forward_parity(detail=True, **c)
    except Exception as e:
        r = dict(error=repr(e), tb=traceback.format_exc())
    r['case'] = c; r['sec'] = time.time() - t0
    print(json.dumps(r, default=str)); sys.stdout.flush()
    out.append(r)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'debug.json'), 'w'), indent=1, default=str)
