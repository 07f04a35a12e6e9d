set -x
TAG=${TAG:-r2g}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_bench.log 2>&1
python scripts/last_step.py gpurun_out/${TAG}_launches.csv | tee gpurun_out/${TAG}_step.txt
