# r3q: pair count variants (CTAs per SM the registers are limited for, entries per step), SDK tests again, whole GPU suite on the new default
TAG=${1:-r3q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_gpu_sdk_selection.py -m gpu -x -q > gpurun_out/pytest_sdk_$TAG.log 2>&1; echo "sdk rc=$? $(tail -1 gpurun_out/pytest_sdk_$TAG.log)"
bash scripts/gpu_ab.sh $TAG d88 "" v98 "HGA_PAIR_OCC=98" v108 "HGA_PAIR_OCC=108" v616 "HGA_PAIR_OCC=616" v816 "HGA_PAIR_OCC=816" v14 "HGA_PAIR_OCC=4"
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "all rc=$? $(tail -1 gpurun_out/pytest_$TAG.log)"
