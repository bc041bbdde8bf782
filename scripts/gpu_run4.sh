# quick A/B: tests + 20/100 Mbp with env variants
TAG=${1:-x}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/pytest_$TAG.log 2>&1; tail -3 gpurun_out/pytest_$TAG.log
run() { # name, env...
  NAME=$1; shift
  env "$@" timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e $GENOME > gpurun_out/bench_${TAG}_$NAME.log 2>&1
  echo "$NAME: $(grep -o '"stages_ms": {[^}]*}' gpurun_out/bench_${TAG}_$NAME.log) $(grep -o '"filter_candidates_per_base": [0-9.]*' gpurun_out/bench_${TAG}_$NAME.log)"
}
GENOME="--genome-mbp 20"
run g20_occ6 HGA_SCAN_MIN_CTAS=6
run g20_occ5 HGA_SCAN_MIN_CTAS=5
run g20_d2 HGA_SCAN_DIAG=2
run g20_d1 HGA_SCAN_DIAG=1
GENOME=""
run g100_occ6 HGA_SCAN_MIN_CTAS=6
run g100_occ5 HGA_SCAN_MIN_CTAS=5
run g100_b8 HGA_FILTER_BITS_PER_KEY=8
run g100_b12 HGA_FILTER_BITS_PER_KEY=12
run g100_b24 HGA_FILTER_BITS_PER_KEY=24 HGA_FILTER_MAX_MB=128
