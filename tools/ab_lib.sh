#!/bin/bash
# A/B of several builds of the library on one GPU box (DIFFPHORE_LIB picks the build; "default" = the in-tree one):
#   gpurun --timeout 600 -- 'bash tools/ab_lib.sh default diffphore_b200/libdiffphore_sm100_base.so ...'
# per build: clock stamps of the second-generation fused convolution at layer 0 (gpurun_out/ab_stamps_<n>.txt) and
# bench.py (headline only); AB_TESTS=1 first runs the kernel-level bit-identity / forward parity tests on the default build.
mkdir -p gpurun_out
if [ -n "$AB_TESTS" ]; then
  timeout 300 python -m pytest tests/test_gpu.py -m gpu -x -q -k 'conv_fused2 or forward_and_update or frozen' 2>&1 | tail -3
fi
n=0
for lib in "$@"; do
  if [ "$lib" = default ]; then export DIFFPHORE_LIB=; else export DIFFPHORE_LIB=$(realpath "$lib"); fi
  timeout 100 python tools/cf2_phases.py 0 > gpurun_out/ab_stamps_$n.txt 2>&1
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --no-e2e > gpurun_out/ab_$n.json 2> gpurun_out/ab_$n.err
  n=$((n+1))
done
python - "$@" <<'PY'
import json, sys
for i, tag in enumerate(sys.argv[1:]):
    try:
        d = json.loads(open(f'gpurun_out/ab_{i}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(tag, 'failed', e, open(f'gpurun_out/ab_{i}.err').read()[-1500:]); continue
    k = d['kernels']
    print(tag.split('/')[-1], round(d['value'], 1), 'samples/s; clocks', d['clocks']['sm_mhz'], '; ' + ' '.join(f"{n.split(':')[1]} {v['ms_per_launch']:.3f}" for n, v in k.items() if n.endswith('0') or n.endswith('tor') or n.endswith('lig3')))
PY
