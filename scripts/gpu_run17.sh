TAG=${1:-x}
mkdir -p gpurun_out
HGA_ENRICH_TIMING=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench100_${TAG}.log 2>&1; echo "bench rc=$?"
grep "hga_enrich:" gpurun_out/bench100_${TAG}.log
