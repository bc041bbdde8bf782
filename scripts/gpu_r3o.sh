# r3o: SDK counting in key-range passes (GPU tests), pivot-subset edge case, then the whole GPU suite
TAG=${1:-r3o}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_gpu_sdk_selection.py -m gpu -x -q > gpurun_out/pytest_sdk_$TAG.log 2>&1; echo "sdk rc=$? $(tail -1 gpurun_out/pytest_sdk_$TAG.log)"
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "all rc=$? $(tail -1 gpurun_out/pytest_$TAG.log)"
