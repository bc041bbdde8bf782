TAG=${1:-r3j}
mkdir -p gpurun_out
bash scripts/gpu_ab.sh $TAG diag2 "HGA_SCAN_DIAG=2" base ""
grep -o '"filter_candidates_per_base": [0-9.]*' gpurun_out/bench_${TAG}_diag2.log gpurun_out/bench_${TAG}_base.log
