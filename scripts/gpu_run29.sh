TAG=${1:-x}
mkdir -p gpurun_out
M=dram__bytes_read.sum,lts__t_sectors.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
for ORD in 1 0; do
HGA_PAIR_ORDER=$ORD timeout 300 ncu --metrics $M --clock-control none -k 'regex:pair_count_warp_kernel' -c 1 --csv --log-file gpurun_out/pairmetrics_${TAG}_ord$ORD.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-enrich > gpurun_out/pairmetrics_${TAG}_ord$ORD.log 2>&1; echo "ord $ORD rc=$?"
done
