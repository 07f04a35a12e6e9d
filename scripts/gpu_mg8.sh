# 8-GPU (or $G) runs of the BASELINE configs quoted on 8 GPUs: c2 (headline), c3 (LFU), c4 (Herald plans, D=512, bound 10)
set -x
mkdir -p gpurun_out
TAG=${TAG:-mg8}
G=${G:-8}
nvidia-smi topo -m | head -14 > gpurun_out/${TAG}_topo.txt
for spec in ${SPECS:-c2 c3 c4:--parity-steps,1}; do
  cfg=${spec%%:*}; extra=""; [ "$spec" != "$cfg" ] && extra=$(echo ${spec#*:} | tr ',' ' ')
  timeout 900 python bench.py --config $cfg --gpus $G --steps 30 --warmup 10 $extra > gpurun_out/${TAG}_${cfg}_g$G.json 2> gpurun_out/${TAG}_${cfg}_g$G.err
  tail -3 gpurun_out/${TAG}_${cfg}_g$G.err
  python - gpurun_out/${TAG}_${cfg}_g$G.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d['config']['workload'],'| N',d['n_gpus'],'ms/step',round(d['ms_per_step'],4),'value',round(d['value']),'e2e',d.get('e2e',{}).get('ms_per_step'))
    print(' phases',d['phases'])
    print(' nvlink',d['roofline'].get('nvlink'))
    p=d.get('parity') or {}
    print(' parity',{k:p.get(k) for k in ('steps','ranks','counters_equal','rows_crc_equal','owner_rows_equal','owner_versions_equal')}, 'planner',d.get('planner'))
except Exception as e: print('ERR',e)
PY
done
