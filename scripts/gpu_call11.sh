set -x
mkdir -p gpurun_out
TAG=${TAG:-c11}
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${TAG}_pytest.log
for tk in 32 16 8; do
HERALD_TICKET_ROWS=$tk timeout 600 python bench.py --steps 50 --warmup 20 --no-cpu-baseline --no-e2e --seg-trace gpurun_out/${TAG}_segtrace_tk$tk.json > gpurun_out/${TAG}_bench_tk$tk.json 2>> gpurun_out/${TAG}_bench.err
done
tail -3 gpurun_out/${TAG}_bench.err
timeout 300 python scripts/laia_bench.py 8 8192 4 3376258 > gpurun_out/${TAG}_laia.txt 2>&1
cat gpurun_out/${TAG}_pytest.log gpurun_out/${TAG}_laia.txt
