TAG=${1:-x}
mkdir -p gpurun_out
(HGA_SCAN_IMPL=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/pytest_${TAG}_impl1.log 2>&1; tail -12 gpurun_out/pytest_${TAG}_impl1.log | cut -c1-300
run() { NAME=$1; shift; env "$@" timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e $GENOME > gpurun_out/bench_${TAG}_$NAME.log 2>&1
  echo "$NAME: $(grep -o '"stages_ms": {[^}]*}' gpurun_out/bench_${TAG}_$NAME.log) $(grep -o '"filter_candidates_per_base": [0-9.]*' gpurun_out/bench_${TAG}_$NAME.log)"; tail -2 gpurun_out/bench_${TAG}_$NAME.log | grep -i "error\|Traceback" ; }
GENOME="--genome-mbp 20"; run g20_impl0 HGA_SCAN_IMPL=0; run g20_impl1 HGA_SCAN_IMPL=1; run g20_impl1_d2 HGA_SCAN_IMPL=1 HGA_SCAN_DIAG=2; run g20_impl1_d1 HGA_SCAN_IMPL=1 HGA_SCAN_DIAG=1
GENOME=""; run g100_impl1 HGA_SCAN_IMPL=1; run g100_impl1_b24 HGA_SCAN_IMPL=1 HGA_FILTER_BITS_PER_KEY=24 HGA_FILTER_MAX_MB=128
