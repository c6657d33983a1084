#!/bin/bash
# Round-2 A/B of the experimental flat weight layout (dp_conv_fused_flat) on a GPU box:
#   gpurun --timeout 900 -- 'bash tools/ab_flat.sh'
# 1. kernel-level bit-identity against the default layout, 2. the whole -m gpu suite on the flat path,
# 3. bench.py with both layouts back to back (same box, same clocks).
set -x
mkdir -p gpurun_out
DIFFPHORE_TEST_FLAT=1 timeout 120 python -m pytest tests/test_gpu.py -m gpu -x -q -k flat_layout 2>&1 | tail -3
# the direct CUDA-vs-reference-output tests written after the round-1 GPU budget ran out (un-gate them once green)
DIFFPHORE_TEST_REFGOLD=1 timeout 120 python -m pytest tests/test_gpu.py -m gpu -q -k 'directly or no_final_step' 2>&1 | tail -5
DIFFPHORE_W2=flat timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 100 python tools/conv_fused_probe.py 2>&1 | tail -8
DIFFPHORE_W2=flat timeout 100 python tools/conv_fused_probe.py 2>&1 | tail -8
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ab_default.json 2> gpurun_out/ab_default.err
DIFFPHORE_W2=flat timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ab_flat.json 2> gpurun_out/ab_flat.err
DIFFPHORE_W2=flat_trim timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
DIFFPHORE_W2=flat_trim timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ab_flat_trim.json 2> gpurun_out/ab_flat_trim.err
python - <<'PY'
import json
for tag in ('default', 'flat', 'flat_trim'):
    d = json.loads(open(f'gpurun_out/ab_{tag}.json').read())
    k = d['kernels']
    print(tag, round(d['value'], 1), 'samples/s; e2e', round(d['e2e']['value'], 1), '; lig3', round(k['conv_fused:lig3']['ms_per_launch'], 3), 'ms; clocks', d['clocks']['sm_mhz'])
PY
