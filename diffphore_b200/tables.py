"""Score-norm tables of the SO(3) and torus diffusions, evaluated lazily for the noise levels actually used.

Product-side mirror of /root/reference/src/utils/so3.py:45-62,92-96 (`_exp_score_norms`, `score_norm`) and
/root/reference/src/utils/torus.py:34-43,75-86 (`score_norm_`, `score_norm`).  The reference builds the full
1000-row / 5001-row tables at import (minutes of numpy, GBs of RAM) although a 20-step schedule touches 20 rows;
here a row is computed on first use (vectorised over the 2000-term series) and cached.

The reference's torus table is an UNSEEDED Monte-Carlo estimate (torus.py:75-79): two reference processes disagree
by ~1 %.  This implementation seeds numpy per row (seed argument) so runs are reproducible.
"""
import numpy as np

MIN_EPS, MAX_EPS, N_EPS, X_N = 0.01, 2, 1000, 2000
X_MIN, TX_N = 1e-5, 5000
SIGMA_MIN, SIGMA_MAX, SIGMA_N = 3e-3, 2, 5000


class So3ScoreNorm:
    def __init__(self):
        self._rows = {}
        self._eps = 10 ** np.linspace(np.log10(MIN_EPS), np.log10(MAX_EPS), N_EPS)
        self._omega = np.linspace(0, np.pi, X_N + 1)[1:]

    @staticmethod
    def index(eps):
        idx = (np.log10(eps) - np.log10(MIN_EPS)) / (np.log10(MAX_EPS) - np.log10(MIN_EPS)) * N_EPS
        return np.clip(np.around(idx).astype(int), a_min=0, a_max=N_EPS - 1)

    def _row(self, i, L=2000, block=250):
        eps, om = self._eps[i], self._omega
        p = np.zeros_like(om)
        ds = np.zeros_like(om)
        lo, dlo = np.sin(om / 2), 0.5 * np.cos(om / 2)
        for l0 in range(0, L, block):
            l = np.arange(l0, min(l0 + block, L))[:, None]
            w = (2 * l + 1) * np.exp(-l * (l + 1) * eps ** 2)
            hi = np.sin(om[None] * (l + 0.5))
            dhi = (l + 0.5) * np.cos(om[None] * (l + 0.5))
            p += (w * hi / lo[None]).sum(0)
            ds += (w * (lo[None] * dhi - hi * dlo[None]) / lo[None] ** 2).sum(0)
        pdf = p * (1 - np.cos(om)) / np.pi
        score = ds / p
        return float(np.sqrt(np.sum(score ** 2 * pdf) / np.sum(pdf) / np.pi))

    def __call__(self, eps):
        eps = np.asarray(eps)
        idx = self.index(eps)
        out = np.empty(idx.shape, dtype=np.float64)
        for k, i in np.ndenumerate(idx):
            if int(i) not in self._rows:
                self._rows[int(i)] = self._row(int(i))
            out[k] = self._rows[int(i)]
        return out

    def score_row(self, i, L=2000, block=250):
        """_score_norms[i, :] of so3.py:35-43,58: d/d omega log of the IGSO(3) series over the omega grid (cached per row)."""
        key = ('score', int(i))
        if key not in self._rows:
            eps, om = self._eps[i], self._omega
            p, ds = np.zeros_like(om), np.zeros_like(om)
            lo, dlo = np.sin(om / 2), 0.5 * np.cos(om / 2)
            for l0 in range(0, L, block):
                l = np.arange(l0, min(l0 + block, L))[:, None]
                w = (2 * l + 1) * np.exp(-l * (l + 1) * eps ** 2)
                hi = np.sin(om[None] * (l + 0.5))
                dhi = (l + 0.5) * np.cos(om[None] * (l + 0.5))
                p += (w * hi / lo[None]).sum(0)
                ds += (w * (lo[None] * dhi - hi * dlo[None]) / lo[None] ** 2).sum(0)
            self._rows[key] = ds / p
        return self._rows[key]

    def score_vec(self, eps, vec):
        """so3.score_vec (so3.py:84-89): score of the rotation vector `vec` at noise level eps."""
        i = int(self.index(np.asarray(eps)))
        om = np.linalg.norm(vec)
        return np.interp(om, self._omega, self.score_row(i)) * vec / om


class TorusScoreNorm:
    def __init__(self, seed=0, n_samples=10000):
        self.seed, self.n_samples, self._rows = seed, n_samples, {}
        self._x = 10 ** np.linspace(np.log10(X_MIN), 0, TX_N + 1) * np.pi
        self._sigma = 10 ** np.linspace(np.log10(SIGMA_MIN), np.log10(SIGMA_MAX), SIGMA_N + 1) * np.pi

    @staticmethod
    def index(sigma):
        s = np.log(sigma / np.pi)
        s = (s - np.log(SIGMA_MIN)) / (np.log(SIGMA_MAX) - np.log(SIGMA_MIN)) * SIGMA_N
        return np.round(np.clip(s, 0, SIGMA_N)).astype(int)

    def score_row(self, i):
        """score_[i, :] = grad / p over the x grid with N = 100 images (torus.py:11-22,38-43); NaN where both underflow."""
        key = ('score', int(i))
        if key not in self._rows:
            x, sig = self._x, self._sigma[i]
            k = np.arange(-100, 101)[:, None]
            xs = x[None] + 2 * np.pi * k
            e = np.exp(-xs ** 2 / 2 / sig ** 2)
            with np.errstate(invalid='ignore', divide='ignore'):
                self._rows[key] = (xs / sig ** 2 * e).sum(0) / e.sum(0)
        return self._rows[key]

    def _row(self, i):
        sig, score_row = self._sigma[i], self.score_row(i)
        rng = np.random.RandomState(self.seed * 100003 + i)
        s = sig * rng.randn(self.n_samples)
        s = (s + np.pi) % (2 * np.pi) - np.pi
        xi = (np.log(np.abs(s) / np.pi) - np.log(X_MIN)) / (0 - np.log(X_MIN)) * TX_N
        xi = np.round(np.clip(xi, 0, TX_N)).astype(int)
        return float(((-np.sign(s) * score_row[xi]) ** 2).mean())

    def __call__(self, sigma):
        sigma = np.asarray(sigma)
        idx = self.index(sigma)
        out = np.empty(idx.shape, dtype=np.float64)
        for k, i in np.ndenumerate(idx):
            if int(i) not in self._rows:
                self._rows[int(i)] = self._row(int(i))
            out[k] = self._rows[int(i)]
        return out

    def score(self, x, sigma):
        """torus.score (torus.py:46-55): table look-up of the wrapped-normal score of angles x at noise level(s) sigma."""
        x = np.asarray(x, dtype=np.float64)
        sigma = np.broadcast_to(np.asarray(sigma, dtype=np.float64), x.shape)
        x = (x + np.pi) % (2 * np.pi) - np.pi
        sign = np.sign(x)
        with np.errstate(divide='ignore'):
            xi = (np.log(np.abs(x) / np.pi) - np.log(X_MIN)) / (0 - np.log(X_MIN)) * TX_N
        xi = np.round(np.clip(xi, 0, TX_N)).astype(int)
        si = self.index(sigma)
        out = np.empty(x.shape, dtype=np.float64)
        for k in np.ndindex(x.shape):
            out[k] = -sign[k] * self.score_row(int(si[k]))[xi[k]]
        return out
