set -x
mkdir -p gpurun_out
TAG=${1:-p}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'segment_reduce_kernel|sort_pass_kernel|sort_hist_all' --launch-skip 24 -c 5 -o gpurun_out/${TAG}_prof -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_prof.log 2>&1
tail -2 gpurun_out/${TAG}_prof.log
