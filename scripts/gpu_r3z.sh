# r3z: index local sort over groups of 16 slots (default now) against groups of 32: parity, A/B
TAG=${1:-r3z}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_properties.py -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest_$TAG.log)"
bash scripts/gpu_ab.sh $TAG l4 "" l5 "HGA_INDEX_LOCAL=5"
