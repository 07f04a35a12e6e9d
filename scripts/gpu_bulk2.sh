set -x
mkdir -p gpurun_out
for r in 4 8 16; do
for g in 2; do
HERALD_BULK_ROWS=$r timeout 600 python bench.py --gpus $g --steps 30 --warmup 10 --no-e2e --no-cpu-baseline --parity-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ROWS=$r N=$g ms/step', round(d['ms_per_step'],4), 'sync', round(d['phases']['pull_ms']['transfer_time'],4), 'gather', round(d['phases']['pull_ms']['copy_time'],4))"
done
done
