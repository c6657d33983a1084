"""ORACLE (test infrastructure).  Restatement of the score-norm tables of
/root/reference/src/utils/so3.py:21-62,92-96 and /root/reference/src/utils/torus.py:11-86, evaluated only
for the table rows a noise schedule actually touches (the reference builds all 1000 / 5001 rows at import).

H1: torus.score_norm_ is an UNSEEDED Monte-Carlo estimate in the reference (torus.py:75-79); here the same
estimator runs under a fixed numpy seed so that oracle and CUDA path share one table.
"""
import numpy as np

MIN_EPS, MAX_EPS, N_EPS, X_N = 0.01, 2, 1000, 2000                      # so3.py:6-7


def so3_eps_index(eps):
    """so3.py:93-95 — evaluated in the dtype of `eps` (float32 in the reference: eps = rot_sigma.cpu().numpy())."""
    eps_idx = (np.log10(eps) - np.log10(MIN_EPS)) / (np.log10(MAX_EPS) - np.log10(MIN_EPS)) * N_EPS
    return np.clip(np.around(eps_idx).astype(int), a_min=0, a_max=N_EPS - 1)


def so3_exp_score_norm_row(idx, L=2000):
    """_exp_score_norms[idx] (so3.py:54-62)."""
    eps = (10 ** np.linspace(np.log10(MIN_EPS), np.log10(MAX_EPS), N_EPS))[idx]
    omega = np.linspace(0, np.pi, X_N + 1)[1:]
    p = 0                                                                 # _expansion (so3.py:21-25)
    for l in range(L):
        p = p + (2 * l + 1) * np.exp(-l * (l + 1) * eps ** 2) * np.sin(omega * (l + 1 / 2)) / np.sin(omega / 2)
    pdf = p * (1 - np.cos(omega)) / np.pi                                 # _density, marginal (so3.py:28-32)
    dS = 0                                                                # _score (so3.py:35-43)
    for l in range(L):
        hi = np.sin(omega * (l + 1 / 2))
        dhi = (l + 1 / 2) * np.cos(omega * (l + 1 / 2))
        lo = np.sin(omega / 2)
        dlo = 1 / 2 * np.cos(omega / 2)
        dS = dS + (2 * l + 1) * np.exp(-l * (l + 1) * eps ** 2) * (lo * dhi - hi * dlo) / lo ** 2
    score = dS / p
    return float(np.sqrt(np.sum(score ** 2 * pdf) / np.sum(pdf) / np.pi))


class So3ScoreNorm:
    """so3.score_norm(eps) (so3.py:92-96) with lazily evaluated rows."""

    def __init__(self):
        self.rows = {}

    def __call__(self, eps):
        eps = np.asarray(eps)
        idx = so3_eps_index(eps)
        for i in np.unique(idx):
            if int(i) not in self.rows:
                self.rows[int(i)] = so3_exp_score_norm_row(int(i))
        return np.asarray([self.rows[int(i)] for i in idx.reshape(-1)], dtype=np.float64).reshape(idx.shape)


X_MIN, TX_N = 1e-5, 5000                                                  # torus.py:25-26
SIGMA_MIN, SIGMA_MAX, SIGMA_N = 3e-3, 2, 5000


def torus_sigma_index(sigma):
    """torus.py:83-85."""
    s = np.log(sigma / np.pi)
    s = (s - np.log(SIGMA_MIN)) / (np.log(SIGMA_MAX) - np.log(SIGMA_MIN)) * SIGMA_N
    return np.round(np.clip(s, 0, SIGMA_N)).astype(int)


def torus_score_row(idx):
    """score_[idx, :] = grad / p with N = 100 images (torus.py:11-22,38-43); pinned on the reference's own p / grad functions
    (tests/golden/tables_ref.npz, tools/make_tables_golden.py)."""
    x = 10 ** np.linspace(np.log10(X_MIN), 0, TX_N + 1) * np.pi
    sig = (10 ** np.linspace(np.log10(SIGMA_MIN), np.log10(SIGMA_MAX), SIGMA_N + 1) * np.pi)[idx]
    p_ = 0
    g_ = 0
    for i in range(-100, 101):
        e = np.exp(-(x + 2 * np.pi * i) ** 2 / 2 / sig ** 2)
        p_ = p_ + e
        g_ = g_ + (x + 2 * np.pi * i) / sig ** 2 * e
    with np.errstate(invalid="ignore", divide="ignore"):
        return g_ / p_           # NaN where both underflow, exactly as the reference table


def torus_score_norm_row(idx, seed=0, n_samples=10000):
    """score_norm_[idx] (torus.py:75-79): mean over 10000 samples x ~ wrapped N(0, sigma_idx) of score(x, sigma)^2,
    with score from the nearest-grid table score_ = grad/p (N=100 images, torus.py:11-43,46-55)."""
    sig = (10 ** np.linspace(np.log10(SIGMA_MIN), np.log10(SIGMA_MAX), SIGMA_N + 1) * np.pi)[idx]
    score_row = torus_score_row(idx)
    rng = np.random.RandomState(seed * 100003 + idx)
    s = sig * rng.randn(n_samples)                                        # sample (torus.py:69-72)
    s = (s + np.pi) % (2 * np.pi) - np.pi
    sign = np.sign(s)
    xi = np.log(np.abs(s) / np.pi)
    xi = (xi - np.log(X_MIN)) / (0 - np.log(X_MIN)) * TX_N
    xi = np.round(np.clip(xi, 0, TX_N)).astype(int)
    sc = -sign * score_row[xi]
    return float((sc ** 2).mean())


class TorusScoreNorm:
    """torus.score_norm(sigma) (torus.py:82-86), seeded (H1)."""

    def __init__(self, seed=0):
        self.seed, self.rows = seed, {}

    def __call__(self, sigma):
        sigma = np.asarray(sigma)
        idx = torus_sigma_index(sigma)
        for i in np.unique(idx):
            if int(i) not in self.rows:
                self.rows[int(i)] = torus_score_norm_row(int(i), self.seed)
        return np.asarray([self.rows[int(i)] for i in idx.reshape(-1)], dtype=np.float64).reshape(idx.shape)
