# r3y: pairs packed into one u64 out of the kernels (keys-only sort + unpack): parity, then A/B against the unpacked path
TAG=${1:-r3y}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "all rc=$? $(tail -1 gpurun_out/pytest_$TAG.log)"
bash scripts/gpu_ab.sh $TAG packed "" unpacked "HGA_PAIRS_UNPACKED=1"
