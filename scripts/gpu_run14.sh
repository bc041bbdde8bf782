TAG=${1:-x}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/pytest_${TAG}.log 2>&1; tail -3 gpurun_out/pytest_${TAG}.log | cut -c1-300
run() { NAME=$1; shift; env "$@" timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e $GENOME > gpurun_out/bench_${TAG}_$NAME.log 2>&1
  echo "$NAME: $(grep -o '"scan": [0-9.]*' gpurun_out/bench_${TAG}_$NAME.log)"; }
GENOME="--genome-mbp 20"; for C in 5 6 7; do run g20_occ$C HGA_SCAN_MIN_CTAS=$C; done
GENOME=""; for C in 5 6 7; do run g100_occ$C HGA_SCAN_MIN_CTAS=$C; done
