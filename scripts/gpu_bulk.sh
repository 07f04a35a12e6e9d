set -x
mkdir -p gpurun_out
export HERALD_SYNC_BULK=1
timeout 900 python -m pytest tests/test_cache_gpu.py tests/test_golden_gpu.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -5
for b in 0 1; do
for g in 1 2; do
HERALD_SYNC_BULK=$b timeout 600 python bench.py --gpus $g --steps 30 --warmup 10 --no-e2e --no-cpu-baseline --parity-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('BULK=$b N=$g ms/step', round(d['ms_per_step'],4), 'sync', round(d['phases']['pull_ms']['transfer_time'],4), 'gather', round(d['phases']['pull_ms']['copy_time'],4))"
done
done
