# r3h: where the DRAM bytes of the scan and pair kernels come from: default, 32 B L2 fetch granularity, scan without key probes
TAG=${1:-r3h}
mkdir -p gpurun_out
B="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-enrich"
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum"
for V in default fetch32 diag1; do
  E=""; [ $V = fetch32 ] && E="HGA_L2_FETCH=32"; [ $V = diag1 ] && E="HGA_SCAN_DIAG=1"
  env $E timeout 600 ncu --metrics $M --clock-control none -k 'regex:scan_probe_kernel' -s 1 -c 1 --csv --log-file gpurun_out/dram_${TAG}_$V.csv python bench.py $B > gpurun_out/dram_${TAG}_$V.log 2>&1
  echo "$V rc=$?"
done
