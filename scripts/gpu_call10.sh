set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 50 --warmup 20 --no-cpu-baseline --no-e2e --seg-trace gpurun_out/c10_segtrace.json > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err
tail -3 gpurun_out/c10_bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus 4 --steps 30 --warmup 10 --no-e2e > gpurun_out/c10_bench_g4.json 2> gpurun_out/c10_bench_g4.err
tail -3 gpurun_out/c10_bench_g4.err
