TAG=${1:-x}
mkdir -p gpurun_out
run() { NAME=$1; shift; env "$@" timeout 150 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-enrich > gpurun_out/bench_${TAG}_$NAME.log 2>&1
  echo "$NAME rc=$?: $(grep -o '"pair_count": [0-9.]*' gpurun_out/bench_${TAG}_$NAME.log) $(grep -o '"pairs": [0-9]*' gpurun_out/bench_${TAG}_$NAME.log)"; }
run longfirst HGA_PAIR_LONG_FIRST=1
run nolong HGA_PAIR_LONG_FIRST=0
