"""ncu target: a few launches of the layer-3 fused convolution (see tools/conv_fused_probe.py)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conv_fused_probe import run
layer = int(sys.argv[1]) if len(sys.argv) > 1 else 3
run(layer, 148 * 16 * 16, 8, mode=1 if layer != 5 else 0, reps=1)
