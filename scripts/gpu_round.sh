set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
nproc; free -g | head -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r1_smoke.log 2>&1
timeout 900 python bench.py --steps 50 --warmup 20 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
tail -3 gpurun_out/r1_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1_ncu_bench.log 2>&1
cat gpurun_out/r1_pytest.log gpurun_out/r1_smoke.log gpurun_out/r1_bench.json
