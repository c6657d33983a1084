#!/bin/bash
# Evidence run of a finished tree on one GPU box (outputs under gpurun_out/, copied to profiles/ by hand):
#   gpurun --timeout 1200 -- 'bash tools/final_run.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu_final.log; cat gpurun_out/pytest_gpu_final.log
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json | head -c 400; echo
timeout 400 python bench.py --impl reference > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extra --no-e2e > /dev/null 2>&1
python tools/ncu_launch_agg.py gpurun_out/launches_final.csv 2>&1 | head -14
