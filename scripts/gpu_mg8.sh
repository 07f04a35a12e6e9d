set -x
mkdir -p gpurun_out
G=${1:-8}
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "all_gpus" 2>&1 | tail -5 > gpurun_out/mg${G}_pytest.log
cat gpurun_out/mg${G}_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus $G --steps 30 --warmup 10 --no-e2e > gpurun_out/mg${G}_bench.json 2> gpurun_out/mg${G}_bench.err
tail -5 gpurun_out/mg${G}_bench.err
grep "^{" gpurun_out/mg${G}_bench.json | cut -c1-400
