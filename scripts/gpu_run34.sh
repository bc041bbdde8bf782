TAG=${1:-x}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/pytest_${TAG}.log 2>&1; tail -2 gpurun_out/pytest_${TAG}.log | cut -c1-300
run() { NAME=$1; shift; env "$@" timeout 150 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-enrich > gpurun_out/bench_${TAG}_$NAME.log 2>&1
  echo "$NAME rc=$?: $(grep -o '"stages_ms": {[^}]*}' gpurun_out/bench_${TAG}_$NAME.log) $(grep -o '"pair_redo_rows": [0-9]*' gpurun_out/bench_${TAG}_$NAME.log) $(grep -o '"pairs": [0-9]*' gpurun_out/bench_${TAG}_$NAME.log)"; }
run a HGA_X=0
run single HGA_PAIR_SINGLE_PASS=1
