"""Build-container-only: pin the score-norm tables (SURVEY a-13) on the reference's OWN functions, committed as
tests/golden/tables_ref.npz.

    python tools/make_tables_golden.py          (needs /root/reference)
The reference modules cannot simply be imported: at import they build all 1000 (so3) / 5001 x 5001 (torus) table rows and
np.save them into a directory that does not exist on the read-only mount.  Their table-building FUNCTIONS are extracted from the
source files instead (ast, unmodified function bodies: so3._expansion / _density / _score, so3.py:21-43; the first torus.p / grad,
torus.py:11-22) and evaluated for the rows a 20-step schedule touches (rot sigma 0.1..1.5, tor sigma 0.0314..3.14):
* so3_idx / so3_exp_score_norm: _exp_score_norms[idx] as so3.py:54-62 computes them;
* torus_idx / torus_score_rows: score_[idx, ::25] = grad / p with N = 100 images (torus.py:38-43), NaNs where both underflow.
The Monte-Carlo estimate torus.score_norm_ itself is unseeded in the reference (H1) and therefore not a golden vector.
"""
import ast
import os
import sys

import numpy as np

REF = '/root/reference/src/utils'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def extract(path, names, first_only=True):
    """exec the (first) definitions of `names` from a reference source file, nothing else of the module."""
    src = open(path).read()
    tree = ast.parse(src)
    ns = {'np': np}
    done = set()
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names and not (first_only and node.name in done):
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, 'exec'), ns)
            done.add(node.name)
    return ns


def main():
    from oracle.tables import so3_eps_index, torus_sigma_index, MIN_EPS, MAX_EPS, N_EPS, X_N, X_MIN, TX_N, SIGMA_MIN, SIGMA_MAX, SIGMA_N
    sched = np.linspace(1, 0, 21)[:-1]
    so3 = extract(os.path.join(REF, 'so3.py'), {'_expansion', '_density', '_score'})
    eps = np.asarray([0.1 ** (1 - t) * 1.5 ** t for t in sched], dtype=np.float32)
    idx = so3_eps_index(eps)
    eps_grid = 10 ** np.linspace(np.log10(MIN_EPS), np.log10(MAX_EPS), N_EPS)
    omegas = np.linspace(0, np.pi, X_N + 1)[1:]
    rows = []
    for i in idx:
        e = so3['_expansion'](omegas, eps_grid[i])
        pdf = so3['_density'](e, omegas, marginal=True)
        sc = so3['_score'](e, omegas, eps_grid[i])
        rows.append(np.sqrt(np.sum(sc ** 2 * pdf) / np.sum(pdf) / np.pi))           # so3.py:62 for one row
    out = {'so3_eps': eps, 'so3_idx': idx, 'so3_exp_score_norm': np.asarray(rows)}

    class _NoBar:                                       # torus.p / grad iterate over tqdm.trange
        @staticmethod
        def trange(*a):
            return range(*a)
    tor = extract(os.path.join(REF, 'torus.py'), {'p', 'grad'})
    tor['tqdm'] = _NoBar
    sig = np.asarray([0.0314 ** (1 - t) * 3.14 ** t for t in sched], dtype=np.float32)
    tidx = torus_sigma_index(sig)
    x = 10 ** np.linspace(np.log10(X_MIN), 0, TX_N + 1) * np.pi
    sgrid = 10 ** np.linspace(np.log10(SIGMA_MIN), np.log10(SIGMA_MAX), SIGMA_N + 1) * np.pi
    with np.errstate(invalid='ignore', divide='ignore'):
        score_rows = np.stack([(tor['grad'](x, sgrid[i], N=100) / tor['p'](x, sgrid[i], N=100))[::25] for i in tidx])
    out.update(torus_sigma=sig, torus_idx=tidx, torus_score_rows=score_rows)

    # ---- so3.score_vec (so3.py:84-89) and torus.score (torus.py:46-55): the reference functions over lazily evaluated table rows
    class _So3Rows:
        def __getitem__(self, i):
            e = so3['_expansion'](omegas, eps_grid[int(i)])
            return so3['_score'](e, omegas, eps_grid[int(i)])
    sv = extract(os.path.join(REF, 'so3.py'), {'score_vec'})
    sv.update(MIN_EPS=MIN_EPS, MAX_EPS=MAX_EPS, N_EPS=N_EPS, _omegas_array=omegas, _score_norms=_So3Rows())
    rng = np.random.RandomState(5)
    vecs = rng.randn(4, 3) * np.asarray([[0.05], [0.4], [1.0], [1.7]])
    veps = np.asarray([0.11, 0.3, 0.8, 1.4])
    out.update(so3_vec=vecs, so3_vec_eps=veps, so3_score_vec=np.stack([sv['score_vec'](float(e), v) for e, v in zip(veps, vecs)]))

    class _TorusTable:
        def __getitem__(self, key):
            si, xi = key
            with np.errstate(invalid='ignore', divide='ignore'):
                return np.asarray([(tor['grad'](x, sgrid[int(a)], N=100) / tor['p'](x, sgrid[int(a)], N=100))[int(b)] for a, b in zip(si, xi)])
    ts = extract(os.path.join(REF, 'torus.py'), {'score'})
    ts.update(X_MIN=X_MIN, X_N=TX_N, SIGMA_MIN=SIGMA_MIN, SIGMA_MAX=SIGMA_MAX, SIGMA_N=SIGMA_N, score_=_TorusTable())
    tx = np.asarray([0.3, -1.0, 2.9, -0.002, 4.0])
    tsg = np.asarray([0.5, 0.05, 2.0, 0.0314, 1.0])
    out.update(torus_x=tx, torus_x_sigma=tsg, torus_score=ts['score'](tx, tsg))
    path = os.path.join(ROOT, 'tests/golden/tables_ref.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
