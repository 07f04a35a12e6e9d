set -x
mkdir -p gpurun_out
for st in 3 10; do
HERALD_HOT_STAGES=$st timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e --seg-trace gpurun_out/c4_segtrace_s$st.json > gpurun_out/c4_bench_s$st.json 2>> gpurun_out/c4_bench.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'segment_reduce_kernel' --launch-skip 4 -c 1 -o gpurun_out/c4_prof -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c4_prof.log 2>&1
tail -2 gpurun_out/c4_prof.log
