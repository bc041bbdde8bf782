TAG=${1:-r2m}
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:pair_count_warp_kernel -s 1 -c 1 -o gpurun_out/pair_${TAG} -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-enrich > gpurun_out/ncu_${TAG}.log 2>&1
echo "ncu rc=$? $(ls -la gpurun_out/pair_${TAG}.ncu-rep | awk '{print $5}')"
