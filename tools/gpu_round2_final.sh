#!/bin/bash
# GPU box, one GPU: the round's final verification and artifacts (tests, smoke, the three bench lines + reference arm, launch list, amortised rate)
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^    \|^$" | tail -8 > gpurun_out/r02t_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02t_smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02t_bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02t_bench_reference.json 2> gpurun_out/bench_ref.err
python bench.py --workload dressing --steps 5 --warmup 3 > gpurun_out/r02t_bench_dressing.json 2> gpurun_out/bench_dressing.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02t_launches_all.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/launch_bench.log 2>&1
python tools/gpu_amortised.py 4096 2 > gpurun_out/r02t_amortised.json 2> gpurun_out/amortised.err
cat gpurun_out/r02t_gpu_tests.log; tail -n 1 gpurun_out/r02t_smoke.log; cut -c1-300 gpurun_out/r02t_bench.json; cut -c1-300 gpurun_out/r02t_bench_reference.json; cut -c1-200 gpurun_out/r02t_bench_dressing.json; cat gpurun_out/r02t_amortised.json; tail -n 2 gpurun_out/amortised.err; wc -l gpurun_out/r02t_launches_all.csv
