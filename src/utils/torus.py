"""utils.torus of the reference (/root/reference/src/utils/torus.py): `score_norm` (:82-86) over the Monte-Carlo
table `score_norm_` (:75-79), evaluated lazily per sigma row and SEEDED (env DIFFPHORE_TORUS_SEED, default 0) —
the reference's table is unseeded, see SURVEY hazard H1."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from diffphore_b200.tables import TorusScoreNorm, SIGMA_MIN, SIGMA_MAX, SIGMA_N   # noqa: E402,F401

_table = TorusScoreNorm(seed=int(os.environ.get('DIFFPHORE_TORUS_SEED', '0')))


def score_norm(sigma):
    return _table(np.asarray(sigma))


def score(x, sigma):
    """Wrapped-normal score of angles x at noise level sigma (reference torus.py:46-55)."""
    return _table.score(x, sigma)
