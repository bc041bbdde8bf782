TAG=${1:-x}
mkdir -p gpurun_out
run() { NAME=$1; shift; env "$@" timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e $GENOME > gpurun_out/bench_${TAG}_$NAME.log 2>&1
  echo "$NAME: $(grep -o '"scan": [0-9.]*' gpurun_out/bench_${TAG}_$NAME.log) $(grep -o '"filter_candidates_per_base": [0-9.]*\|"filter_bytes": [0-9]*' gpurun_out/bench_${TAG}_$NAME.log | tr '\n' ' ')"; }
GENOME=""
for B in 2 4 6 8 12 16; do run d1_b$B HGA_SCAN_DIAG=1 HGA_FILTER_BITS_PER_KEY=$B; done
for B in 6 8 10 12; do run full_b$B HGA_FILTER_BITS_PER_KEY=$B; done
