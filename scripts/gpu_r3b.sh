TAG=${1:-r3b}
mkdir -p gpurun_out
bash scripts/gpu_ab.sh $TAG base ""
grep -o '"pair_redo_rows": [0-9]*\|"pair_mid_rows": [0-9]*' gpurun_out/bench_${TAG}_base.log
