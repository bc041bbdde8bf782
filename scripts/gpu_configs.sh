# throughput of the other BASELINE configs on one GPU (config 4 is the default bench line)
TAG=${1:-r2n}
mkdir -p gpurun_out
for C in 1 2 3 5; do
  timeout 900 python bench.py --config $C --steps 3 --warmup 3 --no-enrich > gpurun_out/bench_${TAG}_config$C.log 2>&1
  echo "config $C rc=$?"; grep -o '"workload": "[^"]*"\|"value": [0-9.]*\|"stages_ms": {[^}]*}\|"frac": [0-9.]*\|"bases": [0-9]*\|"kmers": [0-9]*\|"hits_per_base": [0-9.]*\|"gpu_launches": [0-9]*\|"same_sample": {[^}]*}' gpurun_out/bench_${TAG}_config$C.log | cut -c1-260
done
