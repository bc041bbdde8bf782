# filter residency: scan stage time without key probes (HGA_SCAN_DIAG=1) against the filter size, 10 Gbases
TAG=${1:-r2c}
mkdir -p gpurun_out
B="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-enrich"
run() { NAME=$1; shift; env "$@" timeout 200 python bench.py $B > gpurun_out/bench_${TAG}_$NAME.log 2>&1; echo "$NAME rc=$?: $(grep -o '"scan": [0-9.]*' gpurun_out/bench_${TAG}_$NAME.log) $(grep -o '"filter_candidates_per_base": [0-9.e-]*' gpurun_out/bench_${TAG}_$NAME.log) $(grep -o '"filter_bytes": [0-9]*' gpurun_out/bench_${TAG}_$NAME.log)"; }
for MB in 8 16 24 32 48 64 96; do run d1_mb$MB HGA_SCAN_DIAG=1 HGA_FILTER_BITS_PER_KEY=16 HGA_FILTER_MAX_MB=$MB; done
HGA_SCAN_DIAG=1 timeout 600 ncu --set full --clock-control none -k regex:scan_probe_kernel -s 3 -c 1 -o gpurun_out/scan_${TAG}_d1_10g -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-enrich > gpurun_out/ncu_${TAG}.log 2>&1
HGA_SCAN_DIAG=1 HGA_FILTER_BITS_PER_KEY=16 HGA_FILTER_MAX_MB=32 timeout 600 ncu --set full --clock-control none -k regex:scan_probe_kernel -s 3 -c 1 -o gpurun_out/scan_${TAG}_d1_10g_mb32 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-enrich > gpurun_out/ncu_${TAG}b.log 2>&1
ls -la gpurun_out/*${TAG}*.ncu-rep
