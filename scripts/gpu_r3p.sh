# r3p: SDK counting tests, then the pair count with eight list entries per step (A/B), then the whole GPU suite
TAG=${1:-r3p}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_gpu_sdk_selection.py -m gpu -x -q > gpurun_out/pytest_sdk_$TAG.log 2>&1; echo "sdk rc=$? $(tail -1 gpurun_out/pytest_sdk_$TAG.log)"
bash scripts/gpu_ab.sh $TAG eps4 "" eps8 "HGA_PAIR_OCC=8" eps8r64 "HGA_PAIR_OCC=88"
grep -o '"pairs": [0-9.]*' gpurun_out/bench_${TAG}_eps*.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "all rc=$? $(tail -1 gpurun_out/pytest_$TAG.log)"
