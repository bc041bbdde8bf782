TAG=${1:-r2k}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_properties.py -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1
echo "parity rc=$?: $(tail -3 gpurun_out/pytest_$TAG.log | tr '\n' ' ')"
HGA_FILTER_K=3 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_${TAG}_k3.log 2>&1
echo "parity(3-bit filter) rc=$?: $(tail -3 gpurun_out/pytest_${TAG}_k3.log | tr '\n' ' ')"
B="--steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-enrich"
run() { NAME=$1; shift; env "$@" timeout 200 python bench.py $B $EXTRA > gpurun_out/bench_${TAG}_$NAME.log 2>&1; echo "$NAME rc=$?: $(grep -o '"stages_ms": {[^}]*}' gpurun_out/bench_${TAG}_$NAME.log | cut -c1-60) $(grep -o '"filter_candidates_per_base": [0-9.e-]*' gpurun_out/bench_${TAG}_$NAME.log) $(grep -o '"table_overflow_keys": [0-9]*' gpurun_out/bench_${TAG}_$NAME.log)"; }
EXTRA="" run g100 X=1
EXTRA="" run g100_occ5 HGA_SCAN_OCC=5
EXTRA="" run g100_occ3 HGA_SCAN_OCC=3
EXTRA="" run g100_k3 HGA_FILTER_K=3
EXTRA="" run g100_k3_occ5 HGA_FILTER_K=3 HGA_SCAN_OCC=5
EXTRA="" run g100_d1 HGA_SCAN_DIAG=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_probe_kernel -s 3 -c 1 -o gpurun_out/scan_${TAG} -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-enrich > gpurun_out/ncu_${TAG}.log 2>&1
echo "ncu rc=$?"
