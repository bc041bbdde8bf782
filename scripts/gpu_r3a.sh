# r3a: pair count after the slot pipeline: prefetch experiments and an occupancy sweep (parity of the pair scores first)
TAG=${1:-r3a}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest_$TAG.log)"
bash scripts/gpu_ab.sh $TAG base "" l1 "HGA_PAIR_EXP=1" l1s "HGA_PAIR_EXP=3" sec "HGA_PAIR_EXP=2" c8 "HGA_PAIR_CTAS=8" c7 "HGA_PAIR_CTAS=7" c6 "HGA_PAIR_CTAS=6" c5 "HGA_PAIR_CTAS=5"
