python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^    \|^$" | tail -6 > gpurun_out/r02y_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02y_smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02y_bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02y_bench_reference.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02y_launches_all.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/launch_bench.log 2>&1
cat gpurun_out/r02y_gpu_tests.log; tail -n 1 gpurun_out/r02y_smoke.log; cut -c1-250 gpurun_out/r02y_bench.json; cut -c1-200 gpurun_out/r02y_bench_reference.json
