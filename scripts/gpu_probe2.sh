set -x
mkdir -p gpurun_out
TAG=${1:-p}
timeout 240 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG}_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'segment_reduce_kernel|gather_rows_kernel' --launch-skip 9 -c 2 -o gpurun_out/${TAG}_seg -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_seg.log 2>&1
cat gpurun_out/${TAG}_pytest.log
