TAG=${1:-x}
mkdir -p gpurun_out
run() { NAME=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-enrich > gpurun_out/bench_${TAG}_$NAME.log 2>&1
  echo "$NAME: $(grep -o '"scan": [0-9.]*' gpurun_out/bench_${TAG}_$NAME.log) $(grep -o '"table_overflow_keys": [0-9]*' gpurun_out/bench_${TAG}_$NAME.log) $(grep -o '"table_bytes": [0-9]*' gpurun_out/bench_${TAG}_$NAME.log)"; }
run c4m0 HGA_X=0
run c1m0 HGA_CHAIN_BUCKETS=1
run c1m1 HGA_CHAIN_BUCKETS=1 HGA_PROBE_MODE=1
run c2m1 HGA_CHAIN_BUCKETS=2 HGA_PROBE_MODE=1
run c4m1 HGA_PROBE_MODE=1
(HGA_CHAIN_BUCKETS=1 HGA_PROBE_MODE=1 timeout 600 python -m pytest tests -m gpu -x -q -k "config1_like or long_reads or kat2 or k_sweep or golden" 2>&1 | tail -2)
