mkdir -p gpurun_out
timeout 300 python bench.py --dim 512 --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/big_d512.json 2> gpurun_out/big_d512.err; tail -2 gpurun_out/big_d512.err; cut -c1-200 gpurun_out/big_d512.json
timeout 300 python bench.py --vocab 100000000 --batch 65536 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/big_c5.json 2> gpurun_out/big_c5.err; tail -2 gpurun_out/big_c5.err; cut -c1-200 gpurun_out/big_c5.json
timeout 300 python bench.py --policy lfu --bound 10 --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/big_lfu.json 2> gpurun_out/big_lfu.err; tail -2 gpurun_out/big_lfu.err; cut -c1-200 gpurun_out/big_lfu.json
