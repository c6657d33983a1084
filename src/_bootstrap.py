"""Make the repo root importable when the reference-style entry points are run with cwd=src/ (README.md:39)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
