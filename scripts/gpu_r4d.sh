# r4d: pair kernel instruction trims (one select chain in the probe loop, list membership of a step's entries from one bit mask): parity, bench
TAG=${1:-r4d}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest_$TAG.log)"
bash scripts/gpu_ab.sh $TAG base ""
