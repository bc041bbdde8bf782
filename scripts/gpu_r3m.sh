TAG=${1:-r3m}
mkdir -p gpurun_out
bash scripts/gpu_ab.sh $TAG base "" prefetch "HGA_SCAN_DIAG=4"
