"""GPU-box aid: per-phase clock stamps of dp_conv_fused2 (CTA 0, pairs 2 and 3) on the synthetic probe convolution.
  DIFFPHORE_CONV_GEN=2 python tools/cf2_phases.py [layer]"""
import ctypes, os, sys, torch
os.environ['DIFFPHORE_CONV_GEN'] = '2'
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conv_fused_probe import run, L, dev
layer = int(sys.argv[1]) if len(sys.argv) > 1 else 3
raw = ctypes.CDLL(L.LIB_PATH)
dbg = torch.zeros(5 * 2 * 40 * 3, dtype=torch.int64, device=dev)
raw.dp_debug_set_cf2_probe(ctypes.c_void_p(dbg.data_ptr()))
run(layer, 148 * 16 * 16, 8, mode=int(os.environ.get("PROBE_MODE", "1")), reps=1)
raw.dp_debug_set_cf2_probe(ctypes.c_void_p(0))
d = dbg.cpu().reshape(5, 2, 40, 3)
t0 = int(d[1, 0, 30, 0])
f = lambda x: int(x) - t0 if int(x) else -1
for pr in (0, 1):
    print(f'=== pair {2 + pr}')
    print('worker0: start %d, rs ready %d, x ready %d | loop end %d, part0 staged/reduced %d, pair end %d' % tuple(f(x) for x in list(d[1, pr, 30]) + list(d[1, pr, 31])))
    print('worker4: start %d, rs ready %d, x ready %d | loop end %d, part0 staged/reduced %d, pair end %d' % tuple(f(x) for x in list(d[2, pr, 30]) + list(d[2, pr, 31])))
    for w in (1, 2):
        print(f'worker tile {w - 1} epilogue: tile barrier %d, residual loads issued %d, part0 staged %d | bar %d strad %d part0 reduced %d' % tuple(f(x) for x in list(d[w, pr, 32]) + list(d[w, pr, 33])))
    print('prep: L1 t0 start %d loaded %d done %d | L1 t1 %d %d %d | L2 t0 wait %d got %d done %d | L2 t1 %d %d %d' % tuple(f(x) for k in range(4) for x in d[3, pr, k]))
    print('prep reduce: start %d first part staged %d done %d' % tuple(f(x) for x in d[3, pr, 4]))
    print('copier: start %d window known %d next-iteration %d' % tuple(f(x) for x in d[4, pr, 0]))
    for i in range(26):
        m, w0, w1 = [f(x) for x in d[0, pr, i]], [f(x) for x in d[1, pr, i]], [f(x) for x in d[2, pr, i]]
        if m[0] < 0 and w0[0] < 0:
            continue
        print(f'item/chunk {i:2d}: issuer0 wait {m[0]:6d}->{m[1]:6d} issued {m[2]:6d} | tile0 wait_full {w0[0]:6d}->{w0[1]:6d} drained {w0[2]:6d} | tile1 {w1[0]:6d}->{w1[1]:6d} drained {w1[2]:6d}')
