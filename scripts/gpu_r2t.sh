# r2t: index with three radix passes + index_local_sort_kernel (parity first), A/B against the four-pass sort
TAG=${1:-r2t}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_properties.py -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest_$TAG.log)"
bash scripts/gpu_ab.sh $TAG local1 "HGA_INDEX_LOCAL=1" local0 "HGA_INDEX_LOCAL=0"
