"""GPU-box aid: where the HOST time of a small end-to-end job goes (DenoisingSampler.run on host graphs, poses back on the host).
  python tools/host_profile.py cfg5|cfg3|cfg1   -> wall time per run, device time of the resident loop, cProfile top of one run"""
import cProfile, io, os, pstats, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'src'))
from diffphore_b200.engine import ModelWeights
from diffphore_b200.sampler import DenoisingSampler
from diffphore_b200.synthetic import make_pairs, random_state_dict, real_example_pairs

what = sys.argv[1] if len(sys.argv) > 1 else 'cfg5'
dev = torch.device('cuda:0')
graphs, samples = {'cfg5': (lambda: make_pairs(1, 128, 16), 40), 'cfg3': (lambda: real_example_pairs(15), 40),
                   'cfg1': (lambda: real_example_pairs(12)[11:], 4)}[what]
graphs = graphs()
sampler = DenoisingSampler(ModelWeights(random_state_dict(0), dev), 20)
gen = torch.Generator(device=dev).manual_seed(0)
for k in range(3):
    sampler.run(graphs, samples, generator=gen, pinned=True)
ts = []
for k in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sampler.run(graphs, samples, generator=gen, pinned=True)
    ts.append(1e3 * (time.perf_counter() - t0))
res = sampler.prepare(graphs, samples)
sampler.reset(res, generator=gen); sampler.run_resident(res, generator=gen); torch.cuda.synchronize()
a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
sampler.reset(res, generator=gen); a0.record(); sampler.run_resident(res, generator=gen); a1.record(); torch.cuda.synchronize()
print(f'{what}: run() wall ms {[round(t, 2) for t in ts]}; resident 20-step loop on the device {a0.elapsed_time(a1):.2f} ms')
del res
def phase(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return r, 1e3 * (t1 - t0), 1e3 * (time.perf_counter() - t0)
for k in range(2):
    res, h0, w0 = phase(lambda: sampler.prepare(graphs, samples))
    _, h1, w1 = phase(lambda: sampler.reset(res, generator=gen))
    _, h2, w2 = phase(lambda: sampler.run_resident(res, generator=gen, whole_loop=False))
    _, h3, w3 = phase(lambda: sampler.run_resident(res, generator=gen))
    print(f'phases (host ms / ms until the device is idle): prepare {h0:.2f} / {w0:.2f}, reset {h1:.2f} / {w1:.2f}, eager loop {h2:.2f} / {w2:.2f}, '
          f'default run_resident {h3:.2f} / {w3:.2f}')
    del res
pr = cProfile.Profile(); pr.enable()
sampler.run(graphs, samples, generator=gen, pinned=True)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45); print(s.getvalue()[:9000])
