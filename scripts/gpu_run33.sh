TAG=${1:-x}
mkdir -p gpurun_out
run() { NAME=$1; shift; env "$@" timeout 150 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-enrich > gpurun_out/bench_${TAG}_$NAME.log 2>&1
  echo "$NAME rc=$?: $(grep -o '"pair_count": [0-9.]*' gpurun_out/bench_${TAG}_$NAME.log) $(grep -o '"pair_redo_rows": [0-9]*' gpurun_out/bench_${TAG}_$NAME.log) $(grep -o '"pairs": [0-9]*' gpurun_out/bench_${TAG}_$NAME.log)"; }
run blocked HGA_PAIR_BLOCKED=1
run inter HGA_PAIR_BLOCKED=0
(HGA_PAIR_BLOCKED=1 timeout 300 python -m pytest tests -m gpu -x -q -k "config1_like or long_reads or golden or heavy or pivot" 2>&1 | tail -2)
