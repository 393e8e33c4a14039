#!/bin/bash
# GPU box: the round's verification + artifacts in one call (tests, smoke, bench, reference arm, ncu launch list, ncu full captures).
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round_artifacts.sh r01d'
TAG=${1:-rXX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/gpu_tests_$TAG.log; cat gpurun_out/gpu_tests_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_l.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_pgs -s 60 -c 1 -o gpurun_out/prof_pgs_$TAG -f python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_c.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_narrow -s 60 -c 1 -o gpurun_out/prof_narrow_$TAG -f python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_c2.log 2>&1
python tools/gpu_bathing_bench.py 2>&1 | tail -2 | tee gpurun_out/bathing_$TAG.log
grep "^{" gpurun_out/bench_$TAG.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['config']['envs_over_contact_budget'], {k: round(v,2) for k,v in list(d['roofline']['per_kernel_ms_per_step'].items())})"
