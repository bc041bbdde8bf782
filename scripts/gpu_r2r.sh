TAG=${1:-r2r}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_probe_kernel -s 1 -c 1 -o gpurun_out/scan_${TAG} -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-enrich > gpurun_out/ncu_${TAG}.log 2>&1
echo "ncu rc=$?"
