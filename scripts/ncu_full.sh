# one ncu --set full capture of the step's main kernels (steady state), see B200_PROFILING.md
set -x
mkdir -p gpurun_out
K='regex:segment_hot_kernel|segment_rows_kernel|sync_kernel|gather_rows_kernel|sel_collect|insert_new|resolve_kernel|accumulate'
SKIP=${SKIP:-57}
COUNT=${COUNT:-8}
TAG=${TAG:-r1}
timeout 1200 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip $SKIP -c $COUNT \
  -o gpurun_out/${TAG}_full -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_full.log 2>&1
tail -5 gpurun_out/${TAG}_full.log
ls -la gpurun_out
