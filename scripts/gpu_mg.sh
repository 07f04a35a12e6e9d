set -x
mkdir -p gpurun_out
TAG=${1:-mg}
G=${2:-2}
nvidia-smi -L
nvidia-smi topo -m | head -12
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log
timeout 400 python bench.py --gpus $G --steps 30 --warmup 10 --no-e2e > gpurun_out/${TAG}_bench_g$G.json 2> gpurun_out/${TAG}_bench_g$G.err
tail -5 gpurun_out/${TAG}_bench_g$G.err
cat gpurun_out/${TAG}_bench_g$G.json | cut -c1-1500
