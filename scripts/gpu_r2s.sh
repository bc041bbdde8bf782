# r2s: scan with chunked hit-buffer allocation / batched tickets / new reorder kernel (parity first), then the pivot-order experiments of the pair count
TAG=${1:-r2s}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_properties.py -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest_$TAG.log)"
bash scripts/gpu_ab.sh $TAG order1 "HGA_PAIR_ORDER=1" order2 "HGA_PAIR_ORDER=2" order2s4 "HGA_PAIR_ORDER=2 HGA_PAIR_ORDER_STRIDE=4" true "HGA_BENCH_TRUE_ORDER=1" order0 "HGA_PAIR_ORDER=0"
