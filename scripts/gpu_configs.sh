# BASELINE configs other than the default c2, on ${G:-1} GPU(s)
set -x
mkdir -p gpurun_out
TAG=${TAG:-cfg}
G=${G:-1}
SPECS=${SPECS:-c1 c3 c4:--no-cpu-baseline c5:--no-cpu-baseline}
for spec in $SPECS; do
  cfg=${spec%%:*}; extra=""; [ "$spec" != "$cfg" ] && extra=$(echo ${spec#*:} | tr ',' ' ')
  timeout 900 python bench.py --config $cfg --gpus $G --steps ${STEPS:-30} --warmup 10 $extra > gpurun_out/${TAG}_${cfg}_g$G.json 2> gpurun_out/${TAG}_${cfg}_g$G.err
  tail -3 gpurun_out/${TAG}_${cfg}_g$G.err
  python - gpurun_out/${TAG}_${cfg}_g$G.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d['config']['workload'],'| ms/step',round(d['ms_per_step'],4),'value',round(d['value']),'e2e',d.get('e2e',{}).get('ms_per_step'))
    print(' roofline',round(d['roofline']['frac'],3),{k:round(v['ms'],4) for k,v in d['roofline']['kernels'].items()}, 'parity',d.get('parity',{}).get('counters_equal'), d.get('parity',{}).get('rows_crc_equal'), 'planner',d.get('planner'))
    print(' ',d['phases'])
except Exception as e: print('ERR',e)
PY
done
