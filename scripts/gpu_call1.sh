set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/c1_pytest.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/c1_smoke.log 2>&1
timeout 600 python bench.py --steps 50 --warmup 20 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
tail -3 gpurun_out/c1_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/c1_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c1_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'segment_reduce_kernel|sort_pass_kernel|insert_new|sel_collect' --launch-skip 30 -c 8 -o gpurun_out/c1_prof -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c1_prof.log 2>&1
tail -2 gpurun_out/c1_prof.log
cat gpurun_out/c1_pytest.log gpurun_out/c1_smoke.log gpurun_out/c1_bench.json
