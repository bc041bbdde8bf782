# round 2, first look at the rewritten scan: parity first, then stage times at 2 and 10 Gbases with the filter-size sweep
TAG=${1:-r2a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_properties.py -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1
echo "parity rc=$?: $(tail -3 gpurun_out/pytest_$TAG.log | tr '\n' ' ')"
B="--steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-enrich"
run() { NAME=$1; shift; env "$@" timeout 200 python bench.py $B $EXTRA > gpurun_out/bench_${TAG}_$NAME.log 2>&1; echo "$NAME rc=$?: $(grep -o '"stages_ms": {[^}]*}' gpurun_out/bench_${TAG}_$NAME.log) $(grep -o '"filter_candidates_per_base": [0-9.e-]*' gpurun_out/bench_${TAG}_$NAME.log) $(grep -o '"table_overflow_keys": [0-9]*' gpurun_out/bench_${TAG}_$NAME.log)"; }
EXTRA="--genome-mbp 20" run g20_b12 X=1
EXTRA="" run g100_b12 X=1
EXTRA="" run g100_b8 HGA_FILTER_BITS_PER_KEY=8
EXTRA="" run g100_b16 HGA_FILTER_BITS_PER_KEY=16 HGA_FILTER_MAX_MB=160
EXTRA="" run g100_d1 HGA_SCAN_DIAG=1
EXTRA="" run g100_d2 HGA_SCAN_DIAG=2
EXTRA="" run g100_d3 HGA_SCAN_DIAG=3
