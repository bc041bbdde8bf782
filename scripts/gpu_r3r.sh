TAG=${1:-r3r}
mkdir -p gpurun_out
B="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-enrich"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:pair_count_warp_kernel' -s 1 -c 1 -o gpurun_out/pair_$TAG -f python bench.py $B > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"
