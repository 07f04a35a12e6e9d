set -x
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/c2_pytest.log
HERALD_PDL=0 timeout 600 python bench.py --steps 50 --warmup 20 --no-cpu-baseline --no-e2e > gpurun_out/c2_bench_nopdl.json 2> gpurun_out/c2_bench.err
timeout 600 python bench.py --steps 50 --warmup 20 --no-cpu-baseline --no-e2e --seg-trace gpurun_out/c2_segtrace.json > gpurun_out/c2_bench_pdl.json 2>> gpurun_out/c2_bench.err
tail -3 gpurun_out/c2_bench.err
cat gpurun_out/c2_pytest.log gpurun_out/c2_bench_nopdl.json gpurun_out/c2_bench_pdl.json
