TAG=${1:-x}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/pytest_$TAG.log 2>&1; tail -3 gpurun_out/pytest_$TAG.log
run() { NAME=$1; shift; env "$@" timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e $GENOME > gpurun_out/bench_${TAG}_$NAME.log 2>&1
  echo "$NAME: $(grep -o '"stages_ms": {[^}]*}' gpurun_out/bench_${TAG}_$NAME.log)"; }
GENOME="--genome-mbp 20"; run g20 A=1; run g20_occ5 HGA_SCAN_MIN_CTAS=5; run g20_d2 HGA_SCAN_DIAG=2; run g20_d1 HGA_SCAN_DIAG=1
GENOME=""; run g100 A=1
