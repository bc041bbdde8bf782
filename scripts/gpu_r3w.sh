# r3w: the scan's tile load as a 1-D bulk copy (cp.async.bulk + mbarrier, double buffered) against the plain 16 B loads: parity with it on, then A/B
TAG=${1:-r3w}
mkdir -p gpurun_out
HGA_SCAN_TMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_properties.py -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest (TMA) rc=$? $(tail -1 gpurun_out/pytest_$TAG.log)"
bash scripts/gpu_ab.sh $TAG ldg "" tma "HGA_SCAN_TMA=1"
