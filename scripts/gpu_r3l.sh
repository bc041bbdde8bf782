# r3l: key sectors chosen by the minimizer (the hits of a run share a sector): time and DRAM bytes
TAG=${1:-r3l}
mkdir -p gpurun_out
B="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-enrich"
M="dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum"
env HGA_SECTOR_BY_MIN=1 timeout 600 ncu --metrics $M --clock-control none -k 'regex:scan_probe_kernel' -s 1 -c 1 --csv --log-file gpurun_out/dram_${TAG}_bymin.csv python bench.py $B > gpurun_out/dram_${TAG}_bymin.log 2>&1
bash scripts/gpu_ab.sh $TAG bymin "HGA_SECTOR_BY_MIN=1" bymin_probes "HGA_SECTOR_BY_MIN=1 HGA_SCAN_DIAG=2" bymin_load4 "HGA_SECTOR_BY_MIN=1 HGA_TABLE_LOAD=0.4"
grep -o '"filter_candidates_per_base": [0-9.]*\|"table_overflow_keys": [0-9]*' gpurun_out/bench_${TAG}_bymin*.log
