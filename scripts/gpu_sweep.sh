mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" python bench.py --steps 40 --warmup 15 --no-cpu-baseline --no-e2e > gpurun_out/sw_$tag.json 2>/dev/null; python - <<PY
import json
b=json.load(open('gpurun_out/sw_$tag.json'))
print('$tag', round(b['ms_per_step'],4), round(b['phases']['push_ms']['copy_time'],4), round(b['phases']['pull_ms']['copy_time'],4))
PY
}
run base A=1
run thr32 HERALD_HOT_THRESHOLD=32
run thr128 HERALD_HOT_THRESHOLD=128
run thr256 HERALD_HOT_THRESHOLD=256
run rows2 HERALD_SEG_ROWS=2
run st4 HERALD_HOT_STAGES=4
run tk16 HERALD_TICKET_ROWS=16
