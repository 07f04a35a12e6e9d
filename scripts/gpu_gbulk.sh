set -x
mkdir -p gpurun_out
HERALD_GATHER_BULK=1 timeout 900 python -m pytest tests/test_cache_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -5
for b in 0 1; do
HERALD_GATHER_BULK=$b timeout 600 python bench.py --steps 40 --warmup 10 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('GBULK=$b ms/step', round(d['ms_per_step'],4), 'sync', round(d['phases']['pull_ms']['transfer_time'],4), 'gather', round(d['phases']['pull_ms']['copy_time'],4))"
done
HERALD_GATHER_BULK=1 timeout 600 python bench.py --config c4 --steps 20 --warmup 10 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('GBULK=1 c4 ms/step', round(d['ms_per_step'],4), 'gather', round(d['phases']['pull_ms']['copy_time'],4))"
