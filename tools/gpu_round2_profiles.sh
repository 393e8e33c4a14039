python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02r_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/launch_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 3400 -c 700 --csv --log-file gpurun_out/r02r_launches_dressing.csv python bench.py --workload dressing --steps 1 --warmup 3 > gpurun_out/launch_dress.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:^k_cloth$" -s 58 -c 1 -o gpurun_out/r02r_k_cloth_dressing python bench.py --workload dressing --steps 2 --warmup 3 > gpurun_out/ncu_dress.log 2>&1
python bench.py --workload dressing --impl reference > gpurun_out/r02r_bench_dressing_reference.json 2> gpurun_out/dress_ref.err
python bench.py --workload bedbathing --steps 10 --warmup 3 > gpurun_out/r02r_bench_bedbathing.json 2> gpurun_out/bath.err
tail -2 gpurun_out/dress_ref.err gpurun_out/bath.err gpurun_out/launch_bench.log gpurun_out/launch_dress.log
cut -c1-400 gpurun_out/r02r_bench_dressing_reference.json; cut -c1-900 gpurun_out/r02r_bench_bedbathing.json
wc -l gpurun_out/r02r_launches.csv gpurun_out/r02r_launches_dressing.csv
