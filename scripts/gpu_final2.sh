# round-2 final validation on one GPU: tests, smoke, default bench (+ reference arm), launch list, ncu --set full
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_bench.log 2>&1
python scripts/last_step.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_step.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'segment_reduce_kernel|gather_rows_kernel|sync_bulk_kernel|resolve_kernel|update_tail_kernel|evict_insert_kernel|sel_log_kernel' --launch-skip 30 -c 14 -o gpurun_out/${TAG}_prof -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_prof.log 2>&1
tail -2 gpurun_out/${TAG}_prof.log
cat gpurun_out/${TAG}_pytest.log gpurun_out/${TAG}_smoke.log gpurun_out/${TAG}_step.txt
cut -c1-3000 gpurun_out/${TAG}_bench.json
cut -c1-800 gpurun_out/${TAG}_bench_ref.json
