TAG=${1:-x}
mkdir -p gpurun_out
run() { NAME=$1; shift; env "$@" timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e $GENOME > gpurun_out/bench_${TAG}_$NAME.log 2>&1
  echo "$NAME: $(grep -o '"stages_ms": {[^}]*}' gpurun_out/bench_${TAG}_$NAME.log) $(grep -o '"filter_candidates_per_base": [0-9.]*' gpurun_out/bench_${TAG}_$NAME.log)"; }
GENOME=""; run g100_nopf HGA_SCAN_DIAG=4; run g100_b4 HGA_FILTER_BITS_PER_KEY=4; run g100_d1 HGA_SCAN_DIAG=1; run g100_d2 HGA_SCAN_DIAG=2
for D in 0 4 1; do
HGA_SCAN_DIAG=$D timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors.sum,gpu__time_duration.sum --clock-control none -k regex:scan_probe_kernel -s 1 -c 1 --csv --log-file gpurun_out/dram_${TAG}_d$D.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > /dev/null 2>&1
echo "diag $D:"; grep scan_probe gpurun_out/dram_${TAG}_d$D.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
