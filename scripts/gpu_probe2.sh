set -x
mkdir -p gpurun_out
TAG=${1:-p}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'segment_reduce_kernel' --launch-skip 4 -c 1 -o gpurun_out/${TAG}_seg -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_seg.log 2>&1
tail -2 gpurun_out/${TAG}_seg.log
