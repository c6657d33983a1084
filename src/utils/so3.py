"""utils.so3 of the reference (/root/reference/src/utils/so3.py): only what the denoising path calls, `score_norm`
(:92-96).  The 1000-row `_exp_score_norms` table is evaluated lazily per row (diffphore_b200/tables.py) instead of
at import; no cache files are written."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from diffphore_b200.tables import So3ScoreNorm, MIN_EPS, MAX_EPS, N_EPS   # noqa: E402,F401

_table = So3ScoreNorm()


def score_norm(eps):
    """eps: CPU tensor of rot sigmas -> float32 tensor (same contract as the reference)."""
    return torch.from_numpy(_table(eps.numpy())).float()


def score_vec(eps, vec):
    """Score of a rotation vector at noise level eps (reference so3.py:84-89)."""
    return _table.score_vec(eps, np.asarray(vec))
