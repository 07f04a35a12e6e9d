# A/B of one environment knob on the multi-GPU bench (device-resident numbers only)
# usage: G=2 KNOB=HERALD_APPLY_PER_SM VALUES="8 6 4" bash scripts/gpu_ab_mg.sh
set -x
mkdir -p gpurun_out
TAG=${TAG:-abmg}
G=${G:-2}
for V in ${VALUES}; do
  env ${KNOB}=$V timeout 300 python bench.py --gpus $G --steps 60 --warmup 15 --no-e2e --parity-steps ${PSTEPS:-0} ${BENCH_ARGS} > gpurun_out/${TAG}_$V.json 2> gpurun_out/${TAG}_$V.err
  python - "gpurun_out/${TAG}_$V.json" "$V" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('AB',sys.argv[2],'ms/step',round(d['ms_per_step'],4),{k:round(v*1e3,1) for k,v in d['phases']['pull_ms'].items()}, {k:round(v*1e3,1) for k,v in d['phases']['push_ms'].items()})
except Exception as e: print('ERR',e, open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
