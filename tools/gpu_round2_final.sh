#!/bin/bash
# GPU box, one GPU: the round's final verification and artifacts (tests, smoke, the three bench lines + reference arm, launch lists,
# one full ncu capture of the cloth kernel, amortised rate).  TAG names the files under gpurun_out/.
TAG=${1:-r02w}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^    \|^$" | tail -8 > gpurun_out/${TAG}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/bench_ref.err
python bench.py --workload dressing --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_dressing.json 2> gpurun_out/bench_dressing.err
python bench.py --workload bedbathing --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_bedbathing.json 2> gpurun_out/bench_bedbathing.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches_all.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/launch_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 3400 -c 700 --csv --log-file gpurun_out/${TAG}_launches_dressing.csv python bench.py --workload dressing --steps 1 --warmup 3 --no-cpu > gpurun_out/launch_dress.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:^k_cloth$" -s 58 -c 1 -f -o gpurun_out/${TAG}_k_cloth_dressing python bench.py --workload dressing --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_dress.log 2>&1
python tools/gpu_amortised.py 4096 2 > gpurun_out/${TAG}_amortised.json 2> gpurun_out/amortised.err
cat gpurun_out/${TAG}_gpu_tests.log; tail -n 1 gpurun_out/${TAG}_smoke.log; cut -c1-300 gpurun_out/${TAG}_bench.json; cut -c1-300 gpurun_out/${TAG}_bench_reference.json; cut -c1-200 gpurun_out/${TAG}_bench_dressing.json; cut -c1-200 gpurun_out/${TAG}_bench_bedbathing.json; cat gpurun_out/${TAG}_amortised.json; tail -n 2 gpurun_out/amortised.err; wc -l gpurun_out/${TAG}_launches_all.csv gpurun_out/${TAG}_launches_dressing.csv
