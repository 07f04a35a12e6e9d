# round 2 profile set: launch list + ncu --set full of the row movers
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_bench.log 2>&1
python scripts/last_step.py gpurun_out/${TAG}_launches.csv | tee gpurun_out/${TAG}_step.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'segment_reduce_kernel|gather_rows_kernel|sync_kernel|seg_plan_kernel|resolve_kernel' --launch-skip 40 -c 10 -o gpurun_out/${TAG}_prof -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_prof.log 2>&1
tail -2 gpurun_out/${TAG}_prof.log
ls -la gpurun_out/${TAG}_prof.ncu-rep
